mkdir -p gpurun_out
echo "=== conv_tc + swin tests"; timeout -k 5 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -3
echo "=== bench"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; tail -3 gpurun_out/r02_bench_j.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_j.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== bench head cuda core"; B200_HEAD_CUDA_CORE=1 timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; tail -3 gpurun_out/r02_bench_k.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_k.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
