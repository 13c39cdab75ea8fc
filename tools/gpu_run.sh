python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_stats.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_stats.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_stats2.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_stats2.log
python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/train_breakdown_stats.csv > gpurun_out/bench_train_stats.json 2> gpurun_out/bench_train_stats.err; echo "exit $?"; cut -c1-220 gpurun_out/bench_train_stats.json
HG_BN_STATS_PASS=1 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_nostats.json 2> gpurun_out/bench_train_nostats.err; echo "exit $?"; cut -c1-220 gpurun_out/bench_train_nostats.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_stats.json 2> gpurun_out/bench_infer_stats.err; echo "exit $?"; cut -c1-220 gpurun_out/bench_infer_stats.json
