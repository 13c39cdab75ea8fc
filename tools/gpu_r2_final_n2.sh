# The driver's own N = 2 invocations, each under a short timeout: reference arm, then the default bench line.
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; grep "^{" gpurun_out/bench_ref_n2.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -1 gpurun_out/bench_n2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','e2e','clocks')}); t=d['train']; print({k: t[k] for k in ('value','ms_per_step','allreduce_ms','allreduce','e2e')})"
