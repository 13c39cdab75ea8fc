# Round-1 evidence run (gpurun -- 'bash tools/gpu_profile_r1.sh'): GPU tests, both bench workloads, the reference arm,
# ncu launch lists and full captures.  Results land in gpurun_out/; summarise with tools/ncu_summary.py into profiles/.
# ncu's -k matches the kernel's BASE name (no namespace, no template arguments).
set -x
mkdir -p gpurun_out
K='conv1x1_kernel|conv3x3_kernel|conv3x3_k3_pair_kernel|conv3x3_pair_kernel|conv_gemm_kernel|stem_conv_kernel|stem_pack_kernel|maxpool2x2_kernel|flip_average_kernel|decode_final_kernel|dwconv3x3_kernel'
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1.log 2>&1; echo "smoke exit $?"
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/breakdown_r1.csv > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench exit $?"
python bench.py --workload train --steps 10 --warmup 3 --breakdown gpurun_out/train_breakdown_r1.csv > gpurun_out/bench_train_r1.json 2> gpurun_out/bench_train_r1.err; echo "bench train exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2>&1; echo "ref exit $?"
# stock PyTorch (cuDNN) on the same GPU: the reference network as plain torch ops, fp32 / TF32 / bf16 channels_last
python tools/gpu_stock_torch_baseline.py gpurun_out/stock_torch_b200.json > gpurun_out/stock_torch.log 2>&1; echo "stock torch exit $?"
# the COCO-shaped configuration (C4: 256x192, J=17): inference forward and training step with per-class breakdowns
python tools/c4_time.py --breakdown > gpurun_out/c4_infer_r1.log 2>&1; echo "c4 infer exit $?"
python tools/c4_train_time.py > gpurun_out/c4_train_r1.log 2>&1; echo "c4 train exit $?"
# paired-CTA kernels against the single-CTA / two-kernel paths (bit level) with timings
python tools/pair_check.py single > gpurun_out/pair_single_r1.log 2>&1; HG_CONV3X3_PAIR=1 python tools/pair_check.py pair > gpurun_out/pair_pair_r1.log 2>&1; echo "pair check exit $?"
python tools/k3_check.py > gpurun_out/k3_check_r1.log 2>&1; echo "k3 check exit $?"
# launch list of ONE pass of the inference step (the eager warm-up pass: same kernels as a graph replay)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -s 0 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_r1.log 2>&1; echo "ncu list exit $?"
# launch list of the training step's eager recording pass (everything from pack_weights on)
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 4200 --csv --log-file gpurun_out/launches_train_r1.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_train_r1.log 2>&1; echo "ncu train list exit $?"
# full captures: the dominant kernel (second conv3x3_kernel launch = layer2's 3x3 128->128 at 64x64), the 1x1s around it
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 1 -c 1 -o gpurun_out/prof_r1_conv3x3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1.log 2>&1; echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv1x1_kernel -s 4 -c 6 -o gpurun_out/prof_r1_conv1x1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1_r1.log 2>&1; echo "ncu full 1x1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 0 -c 4 -o gpurun_out/prof_r1_train_wgrad python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_wg.log 2>&1; echo "ncu wgrad exit $?"
ncu --set full --clock-control none --import-source on -k regex:'bn_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_train_fwd_kernel|colstats_kernel' -s 8 -c 8 -o gpurun_out/prof_r1_train_bn python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_bn.log 2>&1; echo "ncu bn exit $?"
tail -3 gpurun_out/pytest_gpu_r1.log; cut -c1-250 gpurun_out/bench_r1.json gpurun_out/bench_train_r1.json
