timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "continuous or groups" > gpurun_out/t_batch2.log 2>&1; echo "rc $?"; tail -n 15 gpurun_out/t_batch2.log
