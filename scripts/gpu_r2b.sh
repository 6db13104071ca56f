#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=3000
timeout 300 python -m pytest tests/test_gpu_device_loop.py -x -q -m gpu > gpurun_out/r2b_pytest_loop.log 2>&1
echo "loop tests rc=$?" | tee gpurun_out/r2b_summary.txt
tail -3 gpurun_out/r2b_pytest_loop.log
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_device_loop.py > gpurun_out/r2b_pytest_all.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/r2b_summary.txt
tail -3 gpurun_out/r2b_pytest_all.log
timeout 120 python scripts/diag_loop5.py 512 > gpurun_out/r2b_diag512.txt 2>&1; echo "diag512 rc=$?" | tee -a gpurun_out/r2b_summary.txt
timeout 120 python scripts/diag_loop5.py 2048 > gpurun_out/r2b_diag2048.txt 2>&1; echo "diag2048 rc=$?" | tee -a gpurun_out/r2b_summary.txt
tail -30 gpurun_out/r2b_diag2048.txt
timeout 200 python scripts/ab_partials.py > gpurun_out/r2b_ab_unit.txt 2>&1; echo "ab unit rc=$?" | tee -a gpurun_out/r2b_summary.txt
TDX_CTA_PARTIALS=1 timeout 200 python scripts/ab_partials.py > gpurun_out/r2b_ab_cta.txt 2>&1; echo "ab cta rc=$?" | tee -a gpurun_out/r2b_summary.txt
paste -d'\n' gpurun_out/r2b_ab_unit.txt gpurun_out/r2b_ab_cta.txt
TDX_LIB=$PWD/tornadox_b200/lib/libtornadox_b200_timeline.so timeout 300 python scripts/attempt_timeline.py > gpurun_out/r2b_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2b_summary.txt
cat gpurun_out/r2b_timeline.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2b_summary.txt
tail -c 600 gpurun_out/r2b_bench.err
