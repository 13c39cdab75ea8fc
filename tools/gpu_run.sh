python -m pytest tests/test_gpu_train.py -x -q -k "depthwise or multi_stream or trajectory" > gpurun_out/pytest_a3.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_a3.log
python -m pytest tests/test_gpu_model.py -q > gpurun_out/pytest_a3_model.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_a3_model.log
