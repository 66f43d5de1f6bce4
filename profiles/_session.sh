mkdir -p gpurun_out
echo "=== swin direct tests"; timeout -k 5 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_unetr.py -q -m gpu --timeout 400 --timeout-method=thread -k "fp32_faithful or unetr or attention_kernels" 2>&1 | tail -25
