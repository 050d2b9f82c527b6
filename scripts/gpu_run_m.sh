#!/bin/bash
# Round-2 GPU job M (1 GPU): the bench line with config 1 in `secondary`, then the whole GPU suite on the final tree.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; tail -c 2500 gpurun_out/r2m_bench_1gpu.json; tail -3 gpurun_out/r2m_bench_1gpu.err
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
cp gpurun_out/gpu_test_errors.json gpurun_out/r2m_gpu_test_errors.json 2>/dev/null
