timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/t_batch.log 2>&1; echo "rc $?"; tail -n 3 gpurun_out/t_batch.log
BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 32 128 256 > gpurun_out/bt_nbuf3.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_nbuf3.log | cut -c1-100
DUALAR_ATTN_TPS=4 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 32 128 > gpurun_out/bt_nbuf3_tps4.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_nbuf3_tps4.log | cut -c1-100
