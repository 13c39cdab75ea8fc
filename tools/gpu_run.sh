nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_infer_n2.json 2> gpurun_out/bench_infer_n2.err; echo "infer n2 exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload train > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 exit $?"
python bench.py --gpus 1 --steps 10 --warmup 3 --workload train --no-cpu-baseline > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; echo "train n1 exit $?"
cut -c1-300 gpurun_out/bench_infer_n2.json gpurun_out/bench_train_n2.json gpurun_out/bench_train_n1.json; tail -5 gpurun_out/bench_train_n2.err
