# HISTORICAL: this session ran on an intermediate state of csrc/attn_tc.cu (between commits e2d23aa and a37a933) in which the variants were
# selectable at run time (B200_ATTN_POLY / B200_ATTN_PT / B200_ATTN_PREFETCH); the committed kernel keeps only the winner.
# anti-phase control (B200_ATTN_PHASE = cycles per 32 keys) and L2 prefetch (B200_ATTN_PREFETCH) on the P-in-TMEM kernel
mkdir -p gpurun_out; rm -f gpurun_out/r02_attention_phase.jsonl
for cfg in "0 0" "0 1" "140 1" "175 1" "210 1" "175 0" "250 1"; do
  set -- $cfg
  echo "=== PT=1 PHASE=$1 PREFETCH=$2"
  B200_ATTN_PT=1 B200_ATTN_PHASE=$1 B200_ATTN_PREFETCH=$2 timeout -k 5 100 python profiles/run_attention.py --batch 8 --iters 7 --only-tc --stages 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    d['phase']=$1; d['prefetch']=$2; d['pt']=1
    print(d['shape'], d['shifted'], d['ms']); open('gpurun_out/r02_attention_phase.jsonl','a').write(json.dumps(d)+'\n')"
done
echo "=== parity PT=1 PHASE=175 PREFETCH=1"; B200_ATTN_PT=1 B200_ATTN_PHASE=175 B200_ATTN_PREFETCH=1 timeout -k 5 200 python -m pytest tests/test_gpu_swin.py -x -q -k "attention_tcgen05 or swin_unetr" 2>&1 | tail -2
echo "=== timeline"; B200_ATTN_PT=1 B200_ATTN_PHASE=175 B200_ATTN_PREFETCH=1 B200_ATTN_TRACE=gpurun_out/attn_trace_phase.bin timeout 100 python profiles/run_attention.py --batch 8 --iters 3 --only-tc --stages 1 > /dev/null 2>&1; python profiles/read_attn_trace.py gpurun_out/attn_trace_phase.bin 6 9
