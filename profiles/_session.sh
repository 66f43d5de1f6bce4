mkdir -p gpurun_out
echo "=== conv tests"; timeout -k 5 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --timeout 120 --timeout-method=thread 2>&1 | tail -4
echo "=== swin tests"; timeout -k 5 600 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 200 --timeout-method=thread -k "swin_unetr or deterministic" 2>&1 | tail -3
echo "=== bench folded"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_p.json 2> gpurun_out/r02_bench_p.err; tail -3 gpurun_out/r02_bench_p.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_p.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
