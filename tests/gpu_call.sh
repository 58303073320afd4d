timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/t_batch.log 2>&1; echo "rc $?"; tail -n 5 gpurun_out/t_batch.log
timeout 900 python bench.py --workload utterances --utterances 1024 --batch 128 > gpurun_out/r2_bench_utt1024_b128_1gpu_async.json 2> gpurun_out/utt_b128.err; echo "rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_utt1024_b128_1gpu_async.json')); print(d['value'], d['seconds_max_rank'], d['total_tokens'])"; tail -n 3 gpurun_out/utt_b128.err
timeout 900 python bench.py --workload utterances --utterances 1024 --batch 256 > gpurun_out/r2_bench_utt1024_b256_1gpu_async.json 2> gpurun_out/utt_b256.err; echo "rc $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_utt1024_b256_1gpu_async.json')); print(d['value'], d['seconds_max_rank'], d['total_tokens'])"; tail -n 3 gpurun_out/utt_b256.err
