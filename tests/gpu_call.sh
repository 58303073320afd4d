timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/t_batch.log 2>&1; echo "rc $?"; tail -n 5 gpurun_out/t_batch.log
BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 64 128 256 > gpurun_out/bt_groups.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_groups.log | cut -c1-150
DUALAR_BATCH_GROUP_SLOTS=64 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 128 256 > gpurun_out/bt_groups64.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_groups64.log | cut -c1-150
DUALAR_BATCH_GROUP_SLOTS=16 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 128 > gpurun_out/bt_groups16.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_groups16.log | cut -c1-150
DUALAR_BATCH_GROUP_SLOTS=128 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 128 > gpurun_out/bt_groups128.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_groups128.log | cut -c1-150
