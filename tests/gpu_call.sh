BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 32 128 256 > gpurun_out/bt_mb3.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_mb3.log | cut -c1-100
DUALAR_TC_STAGES=3 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 128 256 > gpurun_out/bt_mb3_st3.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_mb3_st3.log | cut -c1-100
DUALAR_TC_STAGES=2 BT_SKIP_PREFILL=1 timeout 600 python tests/batch_time.py 128 256 > gpurun_out/bt_mb3_st2.log 2>&1; echo "rc $?"; grep "batched decode" gpurun_out/bt_mb3_st2.log | cut -c1-100
