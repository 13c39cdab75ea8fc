python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_a3t.log 2>&1; echo "exit $?"; tail -30 gpurun_out/pytest_a3t.log
