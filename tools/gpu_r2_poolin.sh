timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "pooled_input or halo_padded" 2>&1 | tail -3
timeout 250 python -m pytest tests/test_gpu_model.py tests/test_gpu_c2_parity.py tests/test_gpu_dropin.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for v in 0 1; do if [ $v = 1 ]; then export HG_NO_POOL_IN=1; fi; echo "NO_POOL_IN=$v"; timeout 100 python bench.py --workload infer --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown gpurun_out/breakdown_poolin$v.csv 2>&1 | grep "^{" | cut -c1-200; done
