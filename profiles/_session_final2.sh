# Last measurement session of round 2 (after the attention rewrite): full GPU suite, smoke, the driver's bench command, artifacts.
mkdir -p gpurun_out
echo "=== full GPU test suite"; timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -4
echo "=== smoke"; timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "=== default bench"; timeout -k 5 600 python bench.py > gpurun_out/r02_bench_default_final2.json 2> gpurun_out/r02_bench_default_final2.err; tail -2 gpurun_out/r02_bench_default_final2.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_default_final2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e'],d['roofline'],d['kernels'],d['clocks']);print({k:(v['ms_per_step'],v.get('lazy')) for k,v in d['secondary'].items()})"
echo "=== attention microbenchmark"; timeout -k 5 100 python profiles/run_attention.py --batch 8 --iters 9 > gpurun_out/r02_attention_times_final2.jsonl 2>&1; cut -c1-160 gpurun_out/r02_attention_times_final2.jsonl
echo "=== breakdown (batch 25)"; timeout -k 5 200 python profiles/run_breakdown.py --batch 25 --reps 3 > gpurun_out/r02_breakdown_b25_final2.txt 2>&1; head -14 gpurun_out/r02_breakdown_b25_final2.txt
echo "=== ncu attention"; timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:window_attention_tc --launch-skip 8 --launch-count 2 -o gpurun_out/r02_attn_final2 -f python profiles/run_ncu_forward.py 8 > gpurun_out/ncu_attn_final2.log 2>&1; tail -1 gpurun_out/ncu_attn_final2.log
