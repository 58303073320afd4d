BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 32 128 > gpurun_out/bt_ks.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_ks.log | cut -c1-100
BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 8 > gpurun_out/bt_ks8.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_ks8.log | cut -c1-100
