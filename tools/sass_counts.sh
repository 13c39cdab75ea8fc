#!/bin/bash
# SASS evidence that the GEMM kernels are tcgen05 / TMA / TMEM code: per object file, the number of
#   UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit)
# and UBLKCP (cp.async.bulk, the stem's raw row segments) instructions in `cuobjdump -sass`.  Usage: tools/sass_counts.sh > profiles/rN_sass_counts.txt
cd "$(dirname "$0")/../hourglass-pose-estimation_b200/build" || exit 1
printf "%-22s %8s %12s %8s %12s %8s %6s %7s %7s\n" object UTCHMMA UTCHMMA.2CTA UTMALDG UTMALDG.2CTA UTMASTG LDTM UTCBAR UBLKCP
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { grep -c "$1" <<<"$s"; }
  printf "%-22s %8d %12d %8d %12d %8d %6d %7d %7d\n" "$o" "$(c 'UTCHMMA')" "$(c 'UTCHMMA.2CTA')" "$(c 'UTMALDG')" "$(c 'UTMALDG.*2CTA')" "$(c 'UTMASTG')" "$(c 'LDTM')" "$(c 'UTCBAR')" "$(c 'UBLKCP')"
done
