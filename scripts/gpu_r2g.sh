#!/bin/bash
# GPU pass G (one B200): StepView constants, batched reduce_units, uniform strip direction with sequential block sums
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
L=$PWD/tornadox_b200/lib
S=gpurun_out/r2g_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2g_pytest_all.log 2>&1
echo "all tests rc=$?" | tee -a $S
tail -3 gpurun_out/r2g_pytest_all.log
timeout 200 python scripts/ab_partials.py > gpurun_out/r2g_ab_unit.txt 2>&1; echo "ab rc=$?" | tee -a $S
TDX_CTA_PARTIALS=1 timeout 200 python scripts/ab_partials.py > gpurun_out/r2g_ab_cta.txt 2>&1
paste -d'\n' gpurun_out/r2g_ab_unit.txt gpurun_out/r2g_ab_cta.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py > gpurun_out/r2g_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a $S
cat gpurun_out/r2g_timeline.txt
timeout 200 python scripts/loop_bench.py > gpurun_out/r2g_loop_bench.txt 2>&1; echo "loop bench rc=$?" | tee -a $S
cat gpurun_out/r2g_loop_bench.txt
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?" | tee -a $S
tail -c 300 gpurun_out/r2g_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ek0_fused_fhn -c 1 -o gpurun_out/r2g_prof_ek0 python scripts/prof_step.py 2048 ek0 3 > gpurun_out/r2g_ncu_ek0.log 2>&1; echo "ncu ek0 rc=$?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dek1_kernel -c 1 -o gpurun_out/r2g_prof_dek1ctl python scripts/prof_step.py 2048 ek1ctl 3 > gpurun_out/r2g_ncu_dek1ctl.log 2>&1; echo "ncu dek1 ctl rc=$?" | tee -a $S
timeout 120 python scripts/fp32_accuracy.py > gpurun_out/r2g_fp32.txt 2>&1; echo "fp32 rc=$?" | tee -a $S
tail -20 gpurun_out/r2g_fp32.txt
