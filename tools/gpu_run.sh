set -x
python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train2.log 2>&1; echo "pytest exit $?"
python bench.py --workload train --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/train_breakdown2.csv > gpurun_out/bench_train2.json 2> gpurun_out/bench_train2.err; echo "bench exit $?"
HG_BN_TWO_PASS=1 python bench.py --workload train --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train2_twopass.json 2> gpurun_out/bench_train2_twopass.err; echo "bench exit $?"
tail -5 gpurun_out/pytest_train2.log; cat gpurun_out/bench_train2.json gpurun_out/bench_train2_twopass.json | cut -c1-400
