# Round-2 check run (gpurun -- 'bash tools/gpu_r2_check.sh'): GPU tests, smoke, the default bench line (inference + train
# sub-record + baselines), the reference arm, and the launch list of smoke() as the driver takes it.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_r2.log
python -m pytest tests/test_gpu_c2_parity.py -m gpu -q -s > gpurun_out/pytest_c2_r2.log 2>&1; echo "c2 parity exit $?"; grep -E "8-stack|passed|failed|Error" gpurun_out/pytest_c2_r2.log | head
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_r2.log
python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/breakdown_r2.csv --train-breakdown gpurun_out/train_breakdown_r2.csv > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r2.err; cut -c1-600 gpurun_out/bench_r2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref_r2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches_r2.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke_r2.log 2>&1; echo "ncu smoke exit $?"
grep -c "hg::" gpurun_out/smoke_launches_r2.csv; grep -o '"[a-z_0-9:<>, ]*kernel[^"]*"' gpurun_out/smoke_launches_r2.csv | sort | uniq -c | sort -rn | head -30
