mkdir -p gpurun_out
echo "=== full GPU test suite"; timeout -k 5 1500 python -m pytest tests -x -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -6
echo "=== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== bench quick"; timeout -k 5 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_q.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['e2e']['value'])"
