T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export HFX_BENCH_WATCHDOG=330
date +%T; $T --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_dbg_bench_2gpu.json 2> gpurun_out/r2_dbg_bench_2gpu.err; date +%T
tail -c 900 gpurun_out/r2_dbg_bench_2gpu.json; grep -v "^W10\|^\*\*\*\|OMP_NUM" gpurun_out/r2_dbg_bench_2gpu.err | head -80
