# Round-1 evidence run (one gpurun call): GPU tests, bench, reference arm, ncu launch list and full capture.
set -x
K='conv1x1_kernel|conv3x3_kernel|conv_gemm_kernel|stem_|maxpool|flip_average|decode_final'
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest exit $?"
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/breakdown_r1.csv > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2>&1; echo "ref exit $?"
# launch list of one graph replay (the eager warm-up pass = 399 matching launches is skipped)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 399 -c 399 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_r1.log 2>&1; echo "ncu list exit $?"
# full capture: first 64x64 3x3 128->128 of the first replay (42 launches of that instantiation per pass), + the 1x1s around it
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_kernel<128>' -s 42 -c 1 -o gpurun_out/prof_r1_conv3x3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1.log 2>&1; echo "ncu full exit $?"
