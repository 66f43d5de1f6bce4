mkdir -p gpurun_out
echo "=== attention tests"; timeout -k 5 400 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread -k "attention or swin_unetr" 2>&1 | tail -3
echo "=== attention microbench"; timeout -k 5 300 python profiles/run_attention.py --batch 8 > gpurun_out/r02_attention_times.jsonl 2>&1; head -4 gpurun_out/r02_attention_times.jsonl
echo "=== default bench"; timeout -k 5 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -3 gpurun_out/r02_bench_default.err; cat gpurun_out/r02_bench_default.json | cut -c1-3000
echo "=== C5 x1"; timeout -k 5 600 python bench.py --steps 3 --warmup 3 --workload swin_c5 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_swin_c5_1gpu.json 2> gpurun_out/r02_bench_swin_c5_1gpu.err; tail -2 gpurun_out/r02_bench_swin_c5_1gpu.err; cut -c1-600 gpurun_out/r02_bench_swin_c5_1gpu.json
