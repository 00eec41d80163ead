for mb in 2 3 4; do HFX_P1_MINB=$mb python tools/order_sweep.py 2e7 1 > gpurun_out/r2_s34_p1_mb$mb.json 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:hdg_col -c 1 -s 2 -o gpurun_out/r2_s34_col python tools/order_sweep.py 2e6 2 > gpurun_out/r2_s34_ncu.log 2>&1
ncu -i gpurun_out/r2_s34_col.ncu-rep --page raw --csv > gpurun_out/r2_s34_col_raw.csv
ncu -i gpurun_out/r2_s34_col.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_s34_col_src.csv
cat gpurun_out/r2_s34_p1_mb*.json
