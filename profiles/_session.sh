mkdir -p gpurun_out
echo "=== stems + stats"; timeout -k 5 300 python -m pytest tests/test_gpu_swin.py -q -m gpu -k "cin1 or deterministic or gemm or layernorm" --timeout 120 --timeout-method=thread 2>&1 | tail -15
echo "=== attention tc"; timeout -k 5 300 python -m pytest tests/test_gpu_swin.py -q -m gpu -k "tcgen05" --timeout 100 --timeout-method=thread 2>&1 | tail -25
echo "=== round2"; timeout -k 5 900 python -m pytest tests/test_gpu_round2.py -q -m gpu --timeout 300 --timeout-method=thread -s 2>&1 | tail -40
echo "=== full suite"; timeout -k 5 1200 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread --deselect tests/test_gpu_round2.py 2>&1 | tail -30
echo "=== bench"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.json
echo "=== library baseline"; timeout -k 5 300 python profiles/run_library_baseline.py --batch 4 > gpurun_out/r02_library_baseline.json 2> gpurun_out/r02_library_baseline.err; tail -2 gpurun_out/r02_library_baseline.err; cat gpurun_out/r02_library_baseline.json
echo "=== blend"; timeout -k 5 200 python profiles/run_blend.py 2>&1 | tail -4
