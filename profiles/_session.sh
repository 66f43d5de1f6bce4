mkdir -p gpurun_out
echo "=== conv_tc + swin tests"; timeout -k 5 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -5
echo "=== bench norm ON LOAD"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; tail -3 gpurun_out/r02_bench_h.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_h.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== bench norm UNFUSED"; B200_NORM_UNFUSED=1 timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; tail -3 gpurun_out/r02_bench_i.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_i.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== ncu conv (norm on load) first 2"; timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"conv3x3x3_tc" --launch-skip 19 --launch-count 2 -f -o gpurun_out/r02_conv_norm2 python profiles/run_ncu_forward.py > gpurun_out/r02_conv_norm2_ncu.log 2>&1; tail -1 gpurun_out/r02_conv_norm2_ncu.log
