python -m pytest tests/test_gpu_trained_accuracy.py tests/test_gpu_model.py -x -q -s > gpurun_out/pytest_acc.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_acc.log
