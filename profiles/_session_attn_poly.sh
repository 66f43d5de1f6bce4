# HISTORICAL: this session ran on an intermediate state of csrc/attn_tc.cu (between commits e2d23aa and a37a933) in which the variants were
# selectable at run time (B200_ATTN_POLY / B200_ATTN_PT / B200_ATTN_PREFETCH); the committed kernel keeps only the winner.
# A/B of the FMA-pipe exponential share in window_attention_tc (B200_ATTN_POLY = 0 | 2 | 3), parity under 3, ncu of 0 and 3.
mkdir -p gpurun_out
for P in 0 2 3; do
  echo "=== B200_ATTN_POLY=$P"; B200_ATTN_POLY=$P timeout -k 5 200 python profiles/run_attention.py --batch 8 --iters 7 --only-tc 2>&1 | tee gpurun_out/r02_attention_poly$P.jsonl | cut -c1-200
done
echo "=== parity with POLY=3"; B200_ATTN_POLY=3 timeout -k 5 300 python -m pytest tests/test_gpu_swin.py -x -q -k "attention_tcgen05 or swin_unetr" 2>&1 | tail -3
for P in 0 3; do
echo "=== ncu POLY=$P"; B200_ATTN_POLY=$P timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:window_attention_tc --launch-skip 8 --launch-count 2 -o gpurun_out/r02_attn_poly$P -f python profiles/run_ncu_forward.py 8 > gpurun_out/ncu_attn_poly$P.log 2>&1; tail -1 gpurun_out/ncu_attn_poly$P.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
