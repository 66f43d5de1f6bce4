mkdir -p gpurun_out
echo "=== swin kernels"; timeout -k 5 400 python -m pytest tests/test_gpu_swin.py -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -5
echo "=== attention microbench"; timeout -k 5 300 python profiles/run_attention.py --batch 8 > gpurun_out/r02_attention_times.jsonl 2>&1; cat gpurun_out/r02_attention_times.jsonl
echo "=== bench (tc attention)"; timeout -k 5 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -3 gpurun_out/r02_bench_f.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_f.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['kernels'])"
echo "=== ncu non-conv kernels"; timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:"conv_cin1_tc|mlp_fused|window_attention_tc|norm_act|head_conv_norm" --launch-skip 33 --launch-count 33 -f -o gpurun_out/r02_swin_kernels python profiles/run_ncu_forward.py > gpurun_out/r02_swin_kernels_ncu.log 2>&1; tail -3 gpurun_out/r02_swin_kernels_ncu.log; ls -la gpurun_out/*.ncu-rep | tail -3
