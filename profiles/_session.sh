mkdir -p gpurun_out
echo "=== full GPU test suite"; timeout -k 5 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -15
