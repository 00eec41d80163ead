N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export HFX_BENCH_WATCHDOG=500
date +%T; $T --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_final_bench_${N}gpu.json 2> gpurun_out/r2_final_bench_${N}gpu.err; echo rc=$?; date +%T
$T --master-port 29542 bench.py --gpus $N --config 4 --steps 3 --warmup 3 > gpurun_out/r2_final_config4_${N}gpu.json 2> gpurun_out/r2_final_config4_${N}gpu.err; echo rc=$?; date +%T
$T --master-port 29543 tools/bench_hex.py --gpus $N --cubes 96 > gpurun_out/r2_final_hex_${N}gpu.json 2> gpurun_out/r2_final_hex_${N}gpu.err; echo rc=$?; date +%T
head -c 260 gpurun_out/r2_final_bench_${N}gpu.json; echo; head -c 260 gpurun_out/r2_final_config4_${N}gpu.json; echo; grep -h "^{" gpurun_out/r2_final_hex_${N}gpu.json | cut -c1-260
grep -v "^W10\|^\*\*\*\|OMP_NUM\|^NCCL" gpurun_out/r2_final_bench_${N}gpu.err gpurun_out/r2_final_config4_${N}gpu.err | head -10
