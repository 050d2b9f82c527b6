#!/bin/bash
# Round-2 GPU job O (1 GPU): final evidence for the restructured stencil kernels: GPU suite, bench line, launch list,
# ncu --set full of both stencil kernels inside a bench step, phase probe, backward phase profile, bulk-copy probe.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cp gpurun_out/gpu_test_errors.json gpurun_out/r2o_gpu_test_errors.json 2>/dev/null
python bench.py > gpurun_out/r2o_bench_1gpu.json 2> gpurun_out/r2o_bench_1gpu.err; head -c 600 gpurun_out/r2o_bench_1gpu.json; echo; tail -3 gpurun_out/r2o_bench_1gpu.err
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2o_ncu_launch.log 2>&1; tail -1 gpurun_out/r2o_ncu_launch.log | head -c 300; echo
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sdf_stencil_bwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2o_sdf_stencil_bwd_tc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2o_ncu_bwd.log 2>&1; tail -1 gpurun_out/r2o_ncu_bwd.log | head -c 300; echo
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sdf_stencil_fwd_tc_kernel -c 1 -f -o gpurun_out/r2o_sdf_stencil_fwd_tc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2o_ncu_fwd.log 2>&1; tail -1 gpurun_out/r2o_ncu_fwd.log | head -c 300; echo
timeout 300 python scripts/stencil_phase_probe.py 2>/dev/null | tail -1 > gpurun_out/r2o_stencil_phase_probe.json; cat gpurun_out/r2o_stencil_phase_probe.json
TF_TC_BWD_PROF=1 timeout 200 python scripts/stencil_phase_probe.py --fwd 0 --bwd 0 --reps 1 2>&1 | grep "bwd prof" | tail -3 > gpurun_out/r2o_bwd_phase_cycles.txt; cat gpurun_out/r2o_bwd_phase_cycles.txt
( for cfg in "32768 2 7 200 0 0 1" "32768 4 7 200 0 0 1" "16384 4 7 200 0 0 1" "8192 8 7 200 0 0 1" "65536 2 7 200 0 0 1" "98304 2 7 200 0 0 1" "32768 2 7 200 2 0 1" "16384 2 7 200 0 0 2" "16384 2 7 200 0 0 4" "8192 2 7 200 0 0 8" "32768 2 7 200 0 0 3"; do timeout 30 tests/probes/_bin/bulk_probe $cfg; done ) > gpurun_out/r2o_bulk_probe.jsonl; cat gpurun_out/r2o_bulk_probe.jsonl | tail -3
python scripts/bench_shape_renderer.py 2>/dev/null | tail -1 > gpurun_out/r2o_bench_shape_renderer.json; head -c 400 gpurun_out/r2o_bench_shape_renderer.json; echo
ls -la gpurun_out | grep r2o | awk '{print $5, $9}'
