#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
L=$PWD/tornadox_b200/lib
timeout 60 python scripts/diag_loop5.py 2048 ek0 16 > gpurun_out/r2f_diag_2048_k16.txt 2>&1; echo "diag main k16 rc=$?" | tee gpurun_out/r2f_summary.txt
tail -3 gpurun_out/r2f_diag_2048_k16.txt
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest_all.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/r2f_summary.txt
tail -3 gpurun_out/r2f_pytest_all.log
timeout 200 python scripts/ab_partials.py > gpurun_out/r2f_ab_unit.txt 2>&1; echo "ab rc=$?" | tee -a gpurun_out/r2f_summary.txt
TDX_LIB=$L/libtornadox_b200_nosums.so timeout 200 python scripts/ab_partials.py > gpurun_out/r2f_ab_nosums.txt 2>&1
TDX_CTA_PARTIALS=1 timeout 200 python scripts/ab_partials.py > gpurun_out/r2f_ab_cta.txt 2>&1
paste -d'\n' gpurun_out/r2f_ab_unit.txt gpurun_out/r2f_ab_nosums.txt gpurun_out/r2f_ab_cta.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py > gpurun_out/r2f_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2f_summary.txt
cat gpurun_out/r2f_timeline.txt
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2f_summary.txt
tail -c 400 gpurun_out/r2f_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ek0_fused_fhn -c 1 -o gpurun_out/r2f_prof_ek0 python scripts/prof_step.py 2048 ek0 3 > gpurun_out/r2f_ncu_ek0.log 2>&1; echo "ncu ek0 rc=$?" | tee -a gpurun_out/r2f_summary.txt
timeout 200 python scripts/loop_bench.py > gpurun_out/r2f_loop_bench.txt 2>&1; echo "loop bench rc=$?" | tee -a gpurun_out/r2f_summary.txt
cat gpurun_out/r2f_loop_bench.txt
