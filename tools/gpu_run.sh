python -m pytest tests/test_gpu_dark.py -x -q > gpurun_out/pytest_dark.log 2>&1; echo "exit $?"; tail -30 gpurun_out/pytest_dark.log
