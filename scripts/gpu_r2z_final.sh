#!/bin/bash
# final single-GPU pass of the round on the final tree: full test suite, smoke, bench lines, ncu launch list + full captures of
# the two dominant kernels (what profiles/traffic.json records)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
S=gpurun_out/r2z_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2z_pytest_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $S
tail -2 gpurun_out/r2z_pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S; tail -1 gpurun_out/r2z_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err; echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --steps 10 --warmup 3 --workload brusselator --no-e2e > gpurun_out/r2z_bench_1gpu_brusselator.json 2> gpurun_out/r2z_bench_1gpu_brusselator.err; echo "bench brusselator rc=$?" | tee -a $S
timeout 200 python scripts/ab_partials.py > gpurun_out/r2z_ab.txt 2>&1; cat gpurun_out/r2z_ab.txt
timeout 200 python scripts/ctl_quick.py > gpurun_out/r2z_ctl_quick.txt 2>&1; cat gpurun_out/r2z_ctl_quick.txt
timeout 200 python scripts/nu_bench.py > gpurun_out/r2z_nu_bench.txt 2>&1; cat gpurun_out/r2z_nu_bench.txt
# profiler passes LAST (numbers printed under ncu are not bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dek1_kernel -c 1 -o gpurun_out/r2z_prof_dek1 -f python scripts/prof_step.py 2048 dek1 3 > gpurun_out/r2z_ncu_dek1.log 2>&1; echo "ncu dek1 rc=$?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ek0_fused_fhn -c 1 -o gpurun_out/r2z_prof_ek0 -f python scripts/prof_step.py 2048 ek0noref 3 > gpurun_out/r2z_ncu_ek0.log 2>&1; echo "ncu ek0 rc=$?" | tee -a $S
