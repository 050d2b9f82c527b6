#!/bin/bash
# Round-2 GPU job Q (1 GPU): final tree: whole GPU suite, smoke(), bench line, launch list, sanitizer passes over the restructured
# tensor-core kernels (stencil forward / backward, X^T Y, dense layer) on small stencil / dense-layer tests.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -3
cp gpurun_out/gpu_test_errors.json gpurun_out/r2q_gpu_test_errors.json 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err; head -c 500 gpurun_out/r2q_bench_1gpu.json; echo; tail -3 gpurun_out/r2q_bench_1gpu.err
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2q_ncu_launch.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py tests/test_tc_gpu.py -q -x -k "stencil_forward or stencil_backward or xty or linear" > gpurun_out/r2q_sanitizer_memcheck_tc.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2q_sanitizer_memcheck_tc.log; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2q_sanitizer_memcheck_tc.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py tests/test_tc_gpu.py -q -x -k "stencil_forward or xty or linear" > gpurun_out/r2q_sanitizer_racecheck_tc.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2q_sanitizer_racecheck_tc.log; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2q_sanitizer_racecheck_tc.log | tail -3
ls -la gpurun_out | grep r2q | awk '{print $5, $9}'
