#!/bin/bash
# 2 GPUs: sharded correctness (incl. bit-equality with the single-GPU sums) and a short sharded bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=2000
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/multigpu_check.py > gpurun_out/r2e_multigpu_check_2.log 2>&1
echo "multigpu_check 2 rc=$?" | tee gpurun_out/r2e_summary.txt
tail -6 gpurun_out/r2e_multigpu_check_2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
echo "bench 2 rc=$?" | tee -a gpurun_out/r2e_summary.txt
tail -c 600 gpurun_out/r2e_bench_2gpu.err
tail -c 1500 gpurun_out/r2e_bench_2gpu.json
