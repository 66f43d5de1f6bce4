mkdir -p gpurun_out
echo "=== swin kernels"; timeout -k 5 400 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -15
echo "=== lazy"; timeout -k 5 300 python -m pytest tests/test_gpu_transforms.py -q -m gpu --timeout 150 --timeout-method=thread -k "lazy" 2>&1 | tail -5
echo "=== attention microbench"; timeout -k 5 300 python profiles/run_attention.py --batch 8 > gpurun_out/r02_attention_times.jsonl 2>&1; cat gpurun_out/r02_attention_times.jsonl
echo "=== breakdown"; timeout -k 5 300 python profiles/run_breakdown.py --batch 4 > gpurun_out/r02_breakdown.txt 2>&1; head -50 gpurun_out/r02_breakdown.txt
echo "=== bench (tc attention)"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -3 gpurun_out/r02_bench_e.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_e.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
