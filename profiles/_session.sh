mkdir -p gpurun_out
echo "=== resample tests"; timeout -k 5 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transforms.py tests/test_gpu_resample.py tests/test_gpu_segresnet.py tests/test_gpu_full_size.py tests/test_gpu_round2.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -8
echo "=== C4 bench"; timeout -k 5 600 python bench.py --workload transforms_c4 --steps 3 --warmup 3 > gpurun_out/r02_bench_transforms_c4.json 2> gpurun_out/r02_bench_transforms_c4.err; tail -2 gpurun_out/r02_bench_transforms_c4.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_transforms_c4.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['lazy'],d['roofline'],d['kernels'])"
echo "=== C4 bench gather"; B200_RESAMPLE_GATHER=1 timeout -k 5 600 python bench.py --workload transforms_c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_transforms_c4_gather.json 2> gpurun_out/r02_bench_transforms_c4_gather.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_transforms_c4_gather.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['lazy'],d['kernels'])"
