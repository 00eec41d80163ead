T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export HFX_BENCH_WATCHDOG=150
date +%T; $T --master-port 29521 bench.py --gpus 2 --cubes 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_dbg_bench_2gpu_small.json 2> gpurun_out/r2_dbg_bench_2gpu_small.err; date +%T
tail -c 600 gpurun_out/r2_dbg_bench_2gpu_small.json; grep -v "^W10\|^\*\*\*\|OMP_NUM" gpurun_out/r2_dbg_bench_2gpu_small.err | head -60
