for cfgs in "X=1" "DUALAR_PDL=0" "DUALAR_TC_KSPLIT=1" "DUALAR_TC_KSPLIT=2" "DUALAR_TC_KSPLIT=4" "DUALAR_TC_STAGES=2" "DUALAR_TC_STAGES=4" "DUALAR_TC_FUSE_NORM=0" "DUALAR_BATCH_NSPLIT=1" "DUALAR_BATCH_NSPLIT=8"; do
  echo "== $cfgs"; env BT_SKIP_PREFILL=1 $cfgs timeout 300 python tests/batch_time.py 32 2>&1 | grep "batched decode" | cut -c1-120
done > gpurun_out/r2_batch_sweep.log 2>&1
cat gpurun_out/r2_batch_sweep.log
