# Round-2 run 2: whole GPU suite (no -x), then compute-sanitizer over tools/sanitize_targets.py
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|8-stack|cosine|arg-max|clearing|losses" gpurun_out/pytest_gpu_r2b.log | cut -c1-400 | head -40
python tools/sanitize_targets.py all > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"; tail -3 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  for what in k3 conv train; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py $what > gpurun_out/sanitize_${tool}_${what}.log 2>&1; echo "$tool $what exit $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|bit-identical|ok" gpurun_out/sanitize_${tool}_${what}.log | tail -4
  done
done
