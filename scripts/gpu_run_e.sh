#!/bin/bash
# Round-2 GPU job E: whole GPU suite after the shader kernels; module-path and material benches.
set -u
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu 2>&1 | tail -60 > gpurun_out/r2e_pytest.log
tail -30 gpurun_out/r2e_pytest.log
python scripts/bench_shape_renderer.py --steps 5 > gpurun_out/r2e_bench_shape_renderer.json 2> gpurun_out/r2e_bench_shape_renderer.err; tail -c 1500 gpurun_out/r2e_bench_shape_renderer.json; tail -3 gpurun_out/r2e_bench_shape_renderer.err
