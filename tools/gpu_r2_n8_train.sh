# Eight-GPU training bench, both exchange modes (gpurun --gpus 8 --timeout 200 -- 'bash tools/gpu_r2_n8_train.sh')
mkdir -p gpurun_out
for mode in 0 1; do
  HG_OVERLAP_AR=$mode timeout 85 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus 8 --steps 20 --warmup 3 --workload train > gpurun_out/bench_train_n8_overlap$mode.json 2> gpurun_out/bench_train_n8_overlap$mode.err; echo "bench n8 overlap=$mode exit $?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train_n8_overlap$mode.json') if l.startswith('{')][-1]); print({k: d[k] for k in ('value','ms_per_step','allreduce','loss_first_last')})"
done
