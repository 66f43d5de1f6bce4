mkdir -p gpurun_out
echo "=== breakdown folded BD=4"; timeout -k 5 300 python profiles/run_breakdown.py --batch 25 --reps 3 2>&1 | grep "conv3x3x3_tc\|total" | head -8
echo "=== breakdown folded BD=2"; B200_RES_BD2=1 timeout -k 5 300 python profiles/run_breakdown.py --batch 25 --reps 3 2>&1 | grep "conv3x3x3_tc\|total" | head -8
echo "=== breakdown unfolded"; B200_RES_UNFOLDED=1 timeout -k 5 300 python profiles/run_breakdown.py --batch 25 --reps 3 2>&1 | grep "conv3x3x3_tc\|total" | head -8
