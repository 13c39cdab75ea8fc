python -m pytest tests/test_gpu_preprocess.py -x -q > gpurun_out/pytest_pre.log 2>&1; echo "exit $?"; tail -30 gpurun_out/pytest_pre.log
