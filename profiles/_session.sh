mkdir -p gpurun_out
echo "=== gemm conv3 shape, batch 25"; timeout -k 5 300 python profiles/run_gemm_tc.py 25 10; timeout -k 5 300 python profiles/run_gemm_tc.py 25 11; timeout -k 5 300 python profiles/run_gemm_tc.py 25 0
echo "=== ncu of the stats gemm"; timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" --launch-skip 5 --launch-count 1 -f -o gpurun_out/r02_gemm_conv3_stats python profiles/run_gemm_tc.py 8 11 > gpurun_out/r02_gemm_conv3_ncu.log 2>&1; tail -1 gpurun_out/r02_gemm_conv3_ncu.log
