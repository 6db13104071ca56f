#!/bin/bash
# 8 GPUs: sharded correctness (bit-equality with the single-GPU sums and step sequences), the cross-GPU cost per attempt,
# and the sharded bench lines (fhn_2d 2048^2 and brusselator 1M)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=3000
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S=gpurun_out/r2p_summary_${N}gpu.txt
: > $S
timeout 500 $TR --master-port 29581 scripts/multigpu_check.py > gpurun_out/r2p_multigpu_check_$N.log 2>&1; echo "multigpu_check $N rc=$?" | tee -a $S
grep MULTIGPU gpurun_out/r2p_multigpu_check_$N.log
[ -n "$SKIP_MICRO" ] || timeout 300 $TR --master-port 29582 scripts/comm_microbench.py > gpurun_out/r2p_comm_microbench_$N.log 2>&1; echo "comm_microbench $N rc=$?" | tee -a $S
grep -E "^(Diagonal|Kronecker)" gpurun_out/r2p_comm_microbench_$N.log
timeout 500 $TR --master-port 29583 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2p_bench_${N}gpu.json 2> gpurun_out/r2p_bench_${N}gpu.err; echo "bench fhn $N rc=$?" | tee -a $S
timeout 500 $TR --master-port 29584 bench.py --gpus $N --steps 10 --warmup 3 --workload brusselator > gpurun_out/r2p_bench_${N}gpu_brusselator.json 2> gpurun_out/r2p_bench_${N}gpu_brusselator.err; echo "bench brusselator $N rc=$?" | tee -a $S
[ -n "$SKIP_F32" ] || timeout 500 $TR --master-port 29585 bench.py --gpus $N --steps 10 --warmup 3 --dtype f32 --no-e2e > gpurun_out/r2p_bench_${N}gpu_f32.json 2> gpurun_out/r2p_bench_${N}gpu_f32.err; echo "bench f32 $N rc=$?" | tee -a $S
python - <<PY
import json
for f in ("r2p_bench_${N}gpu.json", "r2p_bench_${N}gpu_brusselator.json", "r2p_bench_${N}gpu_f32.json"):
    try:
        l = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value %.3e (%s) per_launch %.4f ms device_loop %.4f ms | ek0 %.3e (%s) per_launch %.4f device_loop %.4f | parity %s" % (
            l["value"], l["value_path"], l["per_launch"]["ms_per_step"], l["device_loop"]["ms_per_attempt"],
            l["also"]["ek0"]["value"], l["also"]["ek0"]["path"], l["also"]["ek0"]["per_launch"]["ms_per_step"], l["also"]["ek0"]["device_loop"]["ms_per_attempt"],
            {k: v.get("ok") for k, v in (l.get("parity") or {}).items()}))
    except Exception as e:
        print(f, "unreadable:", e)
PY
