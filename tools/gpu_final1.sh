python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_final_tests.log
python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2_final_sweep.json 2> gpurun_out/r2_final_sweep.err
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_cfg4.json 2> gpurun_out/r2_final_cfg4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_final_ncu_b.log 2>&1
cat gpurun_out/r2_final_tests.log; head -c 600 gpurun_out/r2_final_bench.json; echo; head -c 300 gpurun_out/r2_final_ref.json
