timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "deterministic" > gpurun_out/t_batch4.log 2>&1; echo "rc $?"; tail -n 12 gpurun_out/t_batch4.log
