python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_gpu_dropin.py -x -q > gpurun_out/pytest_pool.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_pool.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_pool.csv > gpurun_out/bench_infer_pool.json 2> gpurun_out/bench_infer_pool.err; echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_infer_pool.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])"
HG_NO_POOL_FUSION=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_nopool.json 2> gpurun_out/bench_infer_nopool.err; echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_infer_nopool.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])"
python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_pool.json 2> gpurun_out/bench_train_pool.err; echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_train_pool.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
HG_NO_POOL_FUSION=1 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_nopool.json 2> gpurun_out/bench_train_nopool.err; echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_train_nopool.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
