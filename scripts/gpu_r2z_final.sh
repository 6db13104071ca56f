#!/bin/bash
# final single-GPU pass of the round: full test suite, bench lines (fp64, fp32, brusselator, ek0 headline), ncu launch list + full
# captures of the two dominant kernels, fp32 accuracy at the BASELINE sizes, in-kernel timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TDX_WAIT_TIMEOUT_MS=1500
L=$PWD/tornadox_b200/lib
S=gpurun_out/r2z_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2z_pytest_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $S
tail -2 gpurun_out/r2z_pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err; echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --steps 10 --warmup 3 --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "bench reference rc=$?" | tee -a $S
timeout 600 python bench.py --steps 10 --warmup 3 --workload brusselator > gpurun_out/r2z_bench_1gpu_brusselator.json 2> gpurun_out/r2z_bench_1gpu_brusselator.err; echo "bench brusselator rc=$?" | tee -a $S
timeout 600 python bench.py --steps 10 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/r2z_bench_1gpu_f32.json 2> gpurun_out/r2z_bench_1gpu_f32.err; echo "bench f32 rc=$?" | tee -a $S
timeout 200 python scripts/ab_partials.py > gpurun_out/r2z_ab.txt 2>&1; cat gpurun_out/r2z_ab.txt
timeout 200 python scripts/ctl_quick.py > gpurun_out/r2z_ctl_quick.txt 2>&1; cat gpurun_out/r2z_ctl_quick.txt
timeout 200 python scripts/loop_bench.py > gpurun_out/r2z_loop_bench.txt 2>&1; cat gpurun_out/r2z_loop_bench.txt
timeout 200 python scripts/fp32_accuracy.py > gpurun_out/r2z_fp32.txt 2>&1; echo "fp32 rc=$?" | tee -a $S; cat gpurun_out/r2z_fp32.txt
timeout 200 python scripts/nu_bench.py > gpurun_out/r2z_nu_bench.txt 2>&1; echo "nu bench rc=$?" | tee -a $S; cat gpurun_out/r2z_nu_bench.txt
TDX_LIB=$L/libtornadox_b200_timeline.so timeout 200 python scripts/attempt_timeline.py > gpurun_out/r2z_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a $S
# profiler passes LAST (numbers printed under ncu are not bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dek1_kernel -c 1 -o gpurun_out/r2z_prof_dek1 python scripts/prof_step.py 2048 dek1 3 > gpurun_out/r2z_ncu_dek1.log 2>&1; echo "ncu dek1 rc=$?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ek0_fused_fhn -c 1 -o gpurun_out/r2z_prof_ek0 python scripts/prof_step.py 2048 ek0noref 3 > gpurun_out/r2z_ncu_ek0.log 2>&1; echo "ncu ek0 rc=$?" | tee -a $S
