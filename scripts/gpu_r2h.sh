#!/bin/bash
# GPU pass H: why are the controller kernels slower per attempt at 8.4 M dimensions?  + row-block alignment experiment for the fused EK0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
L=$PWD/tornadox_b200/lib
S=gpurun_out/r2h_summary.txt
: > $S
timeout 300 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_host_path.py -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "tests rc=$?" | tee -a $S
tail -2 gpurun_out/r2h_pytest.log
timeout 200 python scripts/ctl_quick.py > gpurun_out/r2h_ctl_quick.txt 2>&1; echo "ctl quick rc=$?" | tee -a $S
cat gpurun_out/r2h_ctl_quick.txt
TDX_LIB=$L/libtornadox_b200_ctlldg.so timeout 200 python scripts/ctl_quick.py > gpurun_out/r2h_ctl_quick_ldg.txt 2>&1; echo "ctl quick ldg rc=$?" | tee -a $S
cat gpurun_out/r2h_ctl_quick_ldg.txt
TDX_FORCE_COMM_KERNEL=1 timeout 200 python scripts/ctl_quick.py > gpurun_out/r2h_ctl_quick_comm.txt 2>&1; echo "ctl quick comm rc=$?" | tee -a $S
cat gpurun_out/r2h_ctl_quick_comm.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/dek1_cta_spread.py > gpurun_out/r2h_cta_spread.txt 2>&1; echo "spread rc=$?" | tee -a $S
cat gpurun_out/r2h_cta_spread.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py > gpurun_out/r2h_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a $S
tail -16 gpurun_out/r2h_timeline.txt
for rb in 1 8; do
  TDX_EK0_RB=$rb timeout 200 python scripts/ab_partials.py > gpurun_out/r2h_ab_rb$rb.txt 2>&1
  echo "RB=$rb"; grep ek0 gpurun_out/r2h_ab_rb$rb.txt
done
TDX_EK0_L2_KEEP_MB=0 timeout 200 python scripts/ab_partials.py > gpurun_out/r2h_ab_nohint.txt 2>&1; echo "no L2 hints:"; grep ek0 gpurun_out/r2h_ab_nohint.txt
TDX_EK0_L2_KEEP_MB=40 timeout 200 python scripts/ab_partials.py > gpurun_out/r2h_ab_hint40.txt 2>&1; echo "L2 keep 40 MB:"; grep ek0 gpurun_out/r2h_ab_hint40.txt
