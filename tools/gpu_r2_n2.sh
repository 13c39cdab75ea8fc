# Two-GPU run (gpurun --gpus 2 --timeout 500 -- 'bash tools/gpu_r2_n2.sh'): the NCCL data-parallel test, then the training
# bench at N = 2 with the in-graph bucketed all-reduce (HG_OVERLAP_AR=1) and with one whole-buffer all-reduce (default).
# Every command runs under its own short timeout.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_multi_n2.log 2>&1; echo "pytest exit $?"; grep -E "NCCL world|passed|failed|Error|error" gpurun_out/pytest_multi_n2.log | head
for mode in 1 0; do
  HG_OVERLAP_AR=$mode timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 20 --warmup 3 --workload train > gpurun_out/bench_train_n2_overlap$mode.json 2> gpurun_out/bench_train_n2_overlap$mode.err; echo "bench overlap=$mode exit $?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_n2_overlap$mode.json') if l.startswith('{')][-1]); print({k: d[k] for k in ('value','ms_per_step','allreduce','loss_first_last')})"
done
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
