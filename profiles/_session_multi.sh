# usage: bash profiles/_session_multi.sh N     (run under: gpurun --gpus N -- bash profiles/_session_multi.sh N)
N=$1
mkdir -p gpurun_out
run() { # workload tag steps
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps $3 --warmup 3 --workload $1 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_$2_${N}gpu.json 2> gpurun_out/r02_bench_$2_${N}gpu.err
  tail -2 gpurun_out/r02_bench_$2_${N}gpu.err | cut -c1-300
  python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_$2_${N}gpu.json').read().strip().splitlines()[-1]);print('$2', d['n_gpus'], d['ms_per_step'], d['value'], d.get('e2e'), d.get('parity_vs_single_gpu'), d.get('config'))"
}
if [ "$N" = "2" ]; then
  echo "=== NCCL tests"; timeout -k 5 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -8
fi
echo "=== C3 x$N"; timeout -k 5 600 bash -c "$(declare -f run); N=$N; run swin_c3 swin_c3 5"
echo "=== C5 x$N"; timeout -k 5 600 bash -c "$(declare -f run); N=$N; run swin_c5 swin_c5 3"
