#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python scripts/ab_partials.py 2>&1 | grep ek0
timeout 200 python scripts/ctl_quick.py 2>&1 | grep Kronecker
timeout 300 python -m pytest tests/test_gpu_device_loop.py tests/test_gpu_fullsize.py -q -x -k "ek0 or Kronecker or kronecker" 2>&1 | tail -2
