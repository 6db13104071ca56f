#!/bin/bash
# round-2 GPU pass A: new device-loop tests first, then the whole GPU suite, loop timings, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=3000
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_device_loop.py -x -q -m gpu > gpurun_out/r2a_pytest_loop.log 2>&1
echo "loop tests rc=$?" | tee -a gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_pytest_loop.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_device_loop.py > gpurun_out/r2a_pytest_all.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 600 python scripts/loop_bench.py > gpurun_out/r2a_loop_bench.txt 2>&1
echo "loop bench rc=$?" | tee -a gpurun_out/r2a_summary.txt
cat gpurun_out/r2a_loop_bench.txt
timeout 600 python scripts/config_bench.py > gpurun_out/r2a_config_bench.txt 2>&1
echo "config bench rc=$?" | tee -a gpurun_out/r2a_summary.txt
cat gpurun_out/r2a_config_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" | tee -a gpurun_out/r2a_summary.txt
tail -c 1500 gpurun_out/r2a_bench.json
