mkdir -p gpurun_out
echo "=== fused mlp"; timeout -k 5 300 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread -k "fused_mlp or swin_unetr" 2>&1 | tail -15
echo "=== lazy"; timeout -k 5 300 python -m pytest tests/test_gpu_transforms.py -q -m gpu --timeout 150 --timeout-method=thread -k "lazy" 2>&1 | tail -40
echo "=== resample + round2"; timeout -k 5 600 python -m pytest tests/test_gpu_resample.py tests/test_gpu_round2.py tests/test_gpu_unet.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -15
echo "=== bench (HMMA attention)"; B200_ATTN_HMMA=1 timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_d.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
