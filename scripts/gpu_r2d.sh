#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
L=$PWD/tornadox_b200/lib
timeout 300 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_external_f.py -q -m gpu -x > gpurun_out/r2d_pytest_new.log 2>&1
echo "new tests rc=$?" | tee gpurun_out/r2d_summary.txt
tail -5 gpurun_out/r2d_pytest_new.log
timeout 60 python scripts/diag_loop5.py 2048 ek0 16 > gpurun_out/r2d_diag_2048_k16.txt 2>&1; echo "diag main k16 rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -4 gpurun_out/r2d_diag_2048_k16.txt
TDX_LIB=$L/libtornadox_b200_dbg.so timeout 60 python scripts/diag_loop5.py 2048 ek0 16 > gpurun_out/r2d_diag_2048_k16_dbg.txt 2>&1; echo "diag dbg k16 rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -4 gpurun_out/r2d_diag_2048_k16_dbg.txt
timeout 200 python scripts/ab_partials.py > gpurun_out/r2d_ab_unit.txt 2>&1; echo "ab rc=$?" | tee -a gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_ab_unit.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py > gpurun_out/r2d_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_timeline.txt
timeout 300 python -m pytest tests -q -m gpu --deselect tests/test_gpu_device_loop.py --deselect tests/test_gpu_external_f.py > gpurun_out/r2d_pytest_all.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -3 gpurun_out/r2d_pytest_all.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -c 400 gpurun_out/r2d_bench.err
timeout 200 python scripts/fp32_accuracy.py > gpurun_out/r2d_fp32.txt 2>&1; echo "fp32 rc=$?" | tee -a gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_fp32.txt
