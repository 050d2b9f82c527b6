#!/bin/bash
# Round-2 GPU job N (1 GPU): whole GPU suite + bench line + per-kernel launch list + ncu captures of the two stencil kernels
# after the forward / backward restructuring (A operand in tensor memory, multi-warp W ring, packed fp32 epilogues).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cp gpurun_out/gpu_test_errors.json gpurun_out/r2n_gpu_test_errors.json 2>/dev/null
python bench.py > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; head -c 1800 gpurun_out/r2n_bench_1gpu.json; echo; tail -3 gpurun_out/r2n_bench_1gpu.err
TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2n_ncu_launch.log 2>&1; tail -2 gpurun_out/r2n_ncu_launch.log
for k in sdf_stencil_bwd_tc_kernel sdf_stencil_fwd_tc_kernel; do
  TF_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2n_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2n_ncu_$k.log 2>&1; tail -1 gpurun_out/r2n_ncu_$k.log
done
ls -la gpurun_out | grep r2n
