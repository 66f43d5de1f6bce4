mkdir -p gpurun_out
echo "=== unetr tests"; timeout -k 5 600 python -m pytest tests/test_gpu_unetr.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -25
