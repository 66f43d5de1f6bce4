mkdir -p gpurun_out
echo "=== blend tests"; timeout -k 5 900 python -m pytest tests/test_gpu_sliding_window.py tests/test_gpu_full_size.py tests/test_gpu_round2.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -5
echo "=== blend times"; timeout -k 5 300 python profiles/run_blend.py > gpurun_out/r02_blend_times.jsonl 2>&1; cat gpurun_out/r02_blend_times.jsonl
echo "=== blend times (old lean: no)"; 
