for w in 32 16 8 4; do
HG_HALO_MIN_W=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_halo$w.json 2> gpurun_out/bench_infer_halo$w.err; echo "halo_min_w $w exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_infer_halo$w.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
done
