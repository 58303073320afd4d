timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "gpu rc $?" >> gpurun_out/r2_pytest_gpu.log; tail -n 8 gpurun_out/r2_pytest_gpu.log | cut -c1-400
timeout 600 python tests/batch_time.py 32 > gpurun_out/r2_batch_time.log 2>&1; cat gpurun_out/r2_batch_time.log | cut -c1-230
echo "== no fork"; BT_SKIP_PREFILL=1 DUALAR_BATCH_FORK=0 timeout 300 python tests/batch_time.py 32 2>&1 | grep "batched decode" | cut -c1-100
