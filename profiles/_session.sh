mkdir -p gpurun_out
echo "=== resample tests"; timeout -k 5 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_transforms.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -3
echo "=== resample times (tiled)"; timeout -k 5 300 python profiles/run_resample.py 2>&1 | tail -2
echo "=== resample times (gather)"; B200_RESAMPLE_GATHER=1 timeout -k 5 300 python profiles/run_resample.py 2>&1 | tail -2
