T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export HFX_BENCH_WATCHDOG=400
python -m pytest tests -m gpu -q -k "two_gpus or distributed" 2>&1 | tail -2
date +%T; $T --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err; echo rc=$?; date +%T
$T --master-port 29533 tools/bench_hex.py --gpus 2 > gpurun_out/r2_final_hex_2gpu.json 2> gpurun_out/r2_final_hex_2gpu.err; echo rc=$?; date +%T
head -c 300 gpurun_out/r2_final_bench_2gpu.json; echo; grep -h "^{" gpurun_out/r2_final_hex_2gpu.json | cut -c1-300; grep -v "^W10\|^\*\*\*\|OMP_NUM\|^NCCL" gpurun_out/r2_final_bench_2gpu.err | head -20
