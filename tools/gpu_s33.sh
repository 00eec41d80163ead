python -m pytest tests/test_gpu_col_kernel.py tests/test_gpu_p1_kernel.py -x -q 2>&1 | tail -15 > gpurun_out/r2_s33_tests.log
for nw in 2 4 8; do HFX_COL_NW=$nw python tools/order_sweep.py 2e7 2 > gpurun_out/r2_s33_p2_nw$nw.json 2>&1; done
HFX_COL=0 python tools/order_sweep.py 2e7 2 > gpurun_out/r2_s33_p2_fused.json 2>&1
bash tools/ncu_p1.sh
cat gpurun_out/r2_s33_tests.log gpurun_out/r2_s33_p2_*.json
