# window_attention_tc with P in TMEM and double-buffered Q / K / V: parity, anti-phase delay sweep (B200_ATTN_PHASE), phase timeline
mkdir -p gpurun_out; rm -f gpurun_out/r02_attention_phase_db.jsonl
echo "=== parity (phase 0)"; timeout -k 5 200 python -m pytest tests/test_gpu_swin.py -x -q -k "attention_tcgen05 or swin_unetr" 2>&1 | tail -2
for ph in 0 100 130 160 190 220; do
  echo "=== PHASE=$ph"
  B200_ATTN_PHASE=$ph timeout -k 5 100 python profiles/run_attention.py --batch 8 --iters 9 --only-tc --stages 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    d['phase']=$ph
    print(d['shape'], d['shifted'], d['ms']); open('gpurun_out/r02_attention_phase_db.jsonl','a').write(json.dumps(d)+'\n')"
done
echo "=== parity (phase 160)"; B200_ATTN_PHASE=160 timeout -k 5 200 python -m pytest tests/test_gpu_swin.py -x -q -k "attention_tcgen05 or swin_unetr" 2>&1 | tail -2
for ph in 0 160; do
echo "=== timeline phase $ph"; B200_ATTN_PHASE=$ph B200_ATTN_TRACE=gpurun_out/attn_trace_db$ph.bin timeout 100 python profiles/run_attention.py --batch 8 --iters 3 --only-tc --stages 1 > /dev/null 2>&1; python profiles/read_attn_trace.py gpurun_out/attn_trace_db$ph.bin 6 8
done
