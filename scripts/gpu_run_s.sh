#!/bin/bash
# Round-2 GPU job S (1 GPU): final tree after the tiled dPre layout / kept hidden activations: GPU suite, smoke, bench, launch list,
# ncu of the backward stencil + X^T Y, sanitizer passes over the tensor-core kernel tests.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -3
cp gpurun_out/gpu_test_errors.json gpurun_out/r2s_gpu_test_errors.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2s_bench_1gpu.json 2> gpurun_out/r2s_bench_1gpu.err; head -c 400 gpurun_out/r2s_bench_1gpu.json; echo; tail -3 gpurun_out/r2s_bench_1gpu.err
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2s_ncu_launch.log 2>&1
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sdf_stencil_bwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2s_sdf_stencil_bwd_tc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2s_ncu_bwd.log 2>&1; tail -1 gpurun_out/r2s_ncu_bwd.log | head -c 200; echo
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:xty_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2s_xty_tc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2s_ncu_xty.log 2>&1; tail -1 gpurun_out/r2s_ncu_xty.log | head -c 200; echo
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py tests/test_tc_gpu.py -q -x -k "stencil_forward or stencil_backward or xty or linear" > gpurun_out/r2s_sanitizer_memcheck_tc.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2s_sanitizer_memcheck_tc.log; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2s_sanitizer_memcheck_tc.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py tests/test_tc_gpu.py -q -x -k "stencil_forward or xty or linear" > gpurun_out/r2s_sanitizer_racecheck_tc.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2s_sanitizer_racecheck_tc.log; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2s_sanitizer_racecheck_tc.log | tail -3
timeout 300 python scripts/stencil_phase_probe.py --product 2>/dev/null | tail -1
