DUALAR_TC_FUSE_NORM=2 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 8 > gpurun_out/bt_f2_8.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_f2_8.log | cut -c1-100
DUALAR_TC_FUSE_NORM=2 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 32 128 > gpurun_out/bt_f2.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_f2.log | cut -c1-100
