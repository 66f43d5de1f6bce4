mkdir -p gpurun_out
echo "=== tests"; timeout -k 5 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_swin.py tests/test_gpu_unet.py tests/test_gpu_conv_gather.py -q -m gpu --timeout 200 --timeout-method=thread 2>&1 | tail -5
echo "=== attention microbench"; timeout -k 5 300 python profiles/run_attention.py --batch 8 > gpurun_out/r02_attention_times.jsonl 2>&1; head -4 gpurun_out/r02_attention_times.jsonl
echo "=== bench"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err; tail -3 gpurun_out/r02_bench_m.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_m.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
