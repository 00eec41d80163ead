python -m pytest tests/test_gpu_p1_kernel.py -x -q 2>&1 | tail -3 > gpurun_out/r2_s32_tests.log
ncu --set full --clock-control none --import-source on -k regex:hdg_p1g -c 1 -s 2 -o gpurun_out/r2_s32_p1g python tools/order_sweep.py 2e6 1 > gpurun_out/r2_s32_ncu.log 2>&1
ncu -i gpurun_out/r2_s32_p1g.ncu-rep --page raw --csv > gpurun_out/r2_s32_p1g_raw.csv
ncu -i gpurun_out/r2_s32_p1g.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_s32_p1g_src.csv
cat gpurun_out/r2_s32_tests.log
