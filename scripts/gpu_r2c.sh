#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1000
L=$PWD/tornadox_b200/lib
for s in 1024 2048; do
  TDX_LIB=$L/libtornadox_b200_dbg.so timeout 90 python scripts/diag_loop5.py $s ek0 2 > gpurun_out/r2c_diag_$s.txt 2>&1; echo "diag $s rc=$?" | tee -a gpurun_out/r2c_summary.txt
  tail -8 gpurun_out/r2c_diag_$s.txt
done
TDX_EK0_L2_KEEP_MB=0 TDX_LIB=$L/libtornadox_b200_dbg.so timeout 90 python scripts/diag_loop5.py 2048 ek0 2 > gpurun_out/r2c_diag_2048_nohints.txt 2>&1; echo "diag nohints rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -8 gpurun_out/r2c_diag_2048_nohints.txt
TDX_LIB=$L/libtornadox_b200_dbg.so timeout 90 python scripts/diag_loop5.py 2048 dek1 2 > gpurun_out/r2c_diag_2048_dek1.txt 2>&1; echo "diag dek1 rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -4 gpurun_out/r2c_diag_2048_dek1.txt
export TDX_WAIT_TIMEOUT_MS=3000
timeout 300 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_external_f.py -q -m gpu > gpurun_out/r2c_pytest_new.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -8 gpurun_out/r2c_pytest_new.log
TDX_LIB=$L/libtornadox_b200_ek0timing.so timeout 120 python scripts/ek0_timeline.py 2048 > gpurun_out/r2c_ek0_timeline.txt 2>&1; echo "ek0 timeline rc=$?" | tee -a gpurun_out/r2c_summary.txt
head -8 gpurun_out/r2c_ek0_timeline.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py small > gpurun_out/r2c_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2c_summary.txt
cat gpurun_out/r2c_timeline.txt
timeout 200 python scripts/ab_partials.py > gpurun_out/r2c_ab_unit.txt 2>&1; echo "ab rc=$?" | tee -a gpurun_out/r2c_summary.txt
cat gpurun_out/r2c_ab_unit.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ek0_fused_fhn -c 2 -o gpurun_out/r2c_prof_ek0 python scripts/prof_step.py 2048 ek0 3 > gpurun_out/r2c_ncu_ek0.log 2>&1; echo "ncu ek0 rc=$?" | tee -a gpurun_out/r2c_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dek1_kernel -c 2 -o gpurun_out/r2c_prof_dek1 python scripts/prof_step.py 2048 dek1 3 > gpurun_out/r2c_ncu_dek1.log 2>&1; echo "ncu dek1 rc=$?" | tee -a gpurun_out/r2c_summary.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_step.py > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -5 gpurun_out/r2c_memcheck.log
