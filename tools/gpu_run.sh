python tools/gpu_dag_check.py > gpurun_out/dag_check.log 2>&1; echo "exit $?"; cat gpurun_out/dag_check.log | tail -20
