# A/B timing of kernel variants on ONE box: hourglass-pose-estimation_b200/lib_ab/*.so (built locally with -D switches)
# are swapped in as lib/libhgb200.so one after the other, twice round.
L=hourglass-pose-estimation_b200/lib
cp $L/libhgb200.so /tmp/keep.so
for round in 1 2; do
  for v in hourglass-pose-estimation_b200/lib_ab/*.so; do
    cp $v $L/libhgb200.so
    echo "$(basename $v): $(timeout 120 python tools/k3_time.py 2>&1 | tail -1)"
  done
done
cp /tmp/keep.so $L/libhgb200.so
