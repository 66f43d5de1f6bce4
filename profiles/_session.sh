mkdir -p gpurun_out
echo "=== swin kernels"; timeout -k 5 400 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -8
echo "=== round2 + resample + transforms + unet"; timeout -k 5 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_resample.py tests/test_gpu_transforms.py tests/test_gpu_unet.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -30
echo "=== bench tc attention"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_b.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== bench hmma attention"; B200_ATTN_HMMA=1 timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_b_hmma.json 2> gpurun_out/r02_bench_b_hmma.err; tail -3 gpurun_out/r02_bench_b_hmma.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_b_hmma.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== ncu attention"; timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:window_attention_tc -s 2 -c 2 -o gpurun_out/r02_attn_tc python profiles/run_attention.py --batch 4 --iters 1 --only-tc > gpurun_out/r02_attn_ncu.log 2>&1; tail -3 gpurun_out/r02_attn_ncu.log
echo "=== attention microbench"; timeout -k 5 300 python profiles/run_attention.py --batch 8 > gpurun_out/r02_attention_times.jsonl 2>&1; cat gpurun_out/r02_attention_times.jsonl
