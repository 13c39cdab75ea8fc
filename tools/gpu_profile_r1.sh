# Round-1 evidence run (one gpurun call): GPU tests, both bench workloads, reference arm, ncu launch lists and full captures.
set -x
mkdir -p gpurun_out
K='conv1x1_kernel|conv3x3_kernel|conv_gemm_kernel|stem_conv_kernel|stem_pack_kernel|maxpool2x2_kernel|flip_average|decode_final|dwconv'
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1b.log 2>&1; echo "pytest exit $?"
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/breakdown_r1b.csv > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench exit $?"
python bench.py --workload train --steps 10 --warmup 3 --breakdown gpurun_out/train_breakdown_r1b.csv > gpurun_out/bench_train_r1b.json 2> gpurun_out/bench_train_r1b.err; echo "bench train exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1b.json 2>&1; echo "ref exit $?"
# launch list of ONE graph replay of the inference step (399 launches of our kernels per step; the eager warm-up pass is skipped)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 399 -c 399 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_r1b.log 2>&1; echo "ncu list exit $?"
# full capture of the dominant kernel: the first 3x3 128->128 at 64x64 (layer3's bottleneck) plus the first 1x1s of both shapes
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_kernel<128>' -s 0 -c 1 -o gpurun_out/prof_r1b_conv3x3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1b.log 2>&1; echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:'conv1x1_kernel<256, 0>|conv1x1_kernel<128, 1>' -s 0 -c 4 -o gpurun_out/prof_r1b_conv1x1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1_r1b.log 2>&1; echo "ncu full 1x1 exit $?"
tail -3 gpurun_out/pytest_gpu_r1b.log; cut -c1-250 gpurun_out/bench_r1b.json gpurun_out/bench_train_r1b.json
