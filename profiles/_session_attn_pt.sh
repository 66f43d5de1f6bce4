# HISTORICAL: this session ran on an intermediate state of csrc/attn_tc.cu (between commits e2d23aa and a37a933) in which the variants were
# selectable at run time (B200_ATTN_POLY / B200_ATTN_PT / B200_ATTN_PREFETCH); the committed kernel keeps only the winner.
# P kept in TMEM (B200_ATTN_PT=1): parity, A/B timing, phase timeline
mkdir -p gpurun_out
echo "=== parity with PT=1"; B200_ATTN_PT=1 timeout -k 5 200 python -m pytest tests/test_gpu_swin.py -x -q -k "attention_tcgen05 or swin_unetr" 2>&1 | tail -8
for P in 0 1; do
  echo "=== B200_ATTN_PT=$P"; B200_ATTN_PT=$P timeout -k 5 120 python profiles/run_attention.py --batch 8 --iters 7 --only-tc 2>&1 | tee gpurun_out/r02_attention_pt$P.jsonl | cut -c1-175
done
echo "=== timeline PT=1"; B200_ATTN_PT=1 B200_ATTN_TRACE=gpurun_out/attn_trace_pt.bin timeout 100 python profiles/run_attention.py --batch 8 --iters 3 --only-tc --stages 1 > /dev/null 2>&1; python profiles/read_attn_trace.py gpurun_out/attn_trace_pt.bin 6 9
