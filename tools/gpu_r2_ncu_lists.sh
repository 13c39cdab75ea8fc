# Round-2 ncu launch lists (gpurun -- 'bash tools/gpu_r2_ncu_lists.sh'): ONE pass of the inference step (the eager warm-up
# pass: the same kernels as a graph replay) and the training step's eager recording pass.  Summarise with tools/ncu_summary.py.
mkdir -p gpurun_out
K='conv1x1_kernel|conv3x3_kernel|conv3x3_k3_pair_kernel|conv3x3_pair_kernel|conv_gemm_kernel|stem_conv_kernel|stem_pack_kernel|maxpool2x2_kernel|flip_average_kernel|decode_final_kernel|dwconv3x3_kernel'
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 0 -c 304 --csv --log-file gpurun_out/launches_r2.csv python bench.py --workload infer --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_list_r2.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 3000 --csv --log-file gpurun_out/launches_train_r2.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_list_train_r2.log 2>&1; echo "ncu train list exit $?"
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_c2_parity.py tests/test_gpu_dropin.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 100 python bench.py --workload infer --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>&1 | grep "^{" | cut -c1-200
