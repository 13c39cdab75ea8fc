# A/B of 1x1-kernel variants on ONE box: lib_ab/*.so swapped in as lib/libhgb200.so; standalone shapes + the C2 step.
L=hourglass-pose-estimation_b200/lib
cp $L/libhgb200.so /tmp/keep.so
for round in 1 2; do
  for v in hourglass-pose-estimation_b200/lib_ab/*.so; do
    cp $v $L/libhgb200.so
    echo "$(basename $v): $(timeout 100 python tools/k1_profile.py 2>&1 | head -1) | step $(timeout 100 python bench.py --workload infer --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c 'import json,sys; print(json.loads(sys.stdin.read())["ms_per_step"])')"
  done
done
cp /tmp/keep.so $L/libhgb200.so
