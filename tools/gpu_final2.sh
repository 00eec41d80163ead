T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests -m gpu -q -k "two_gpus or distributed" 2>&1 | tail -3 > gpurun_out/r2_final_tests_2gpu.log
$T --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err
$T --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 3 > gpurun_out/r2_final_ref_2gpu.json 2> gpurun_out/r2_final_ref_2gpu.err
python tools/bench_hex.py > gpurun_out/r2_final_hex_1gpu.json 2> gpurun_out/r2_final_hex_1gpu.err
$T --master-port 29513 tools/bench_hex.py --gpus 2 > gpurun_out/r2_final_hex_2gpu.json 2> gpurun_out/r2_final_hex_2gpu.err
cat gpurun_out/r2_final_tests_2gpu.log; head -c 400 gpurun_out/r2_final_bench_2gpu.json; echo; grep -h "^{" gpurun_out/r2_final_hex_1gpu.json gpurun_out/r2_final_hex_2gpu.json | cut -c1-700
