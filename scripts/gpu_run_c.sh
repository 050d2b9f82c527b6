#!/bin/bash
# Round-2 GPU job C: GPU suite after the material-stage fusions, material bench.
set -u
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log
python scripts/bench_material.py --steps 5 > gpurun_out/r2c_bench_material.json 2> gpurun_out/r2c_bench_material.err; tail -c 1500 gpurun_out/r2c_bench_material.json; tail -3 gpurun_out/r2c_bench_material.err
python scripts/bench_shape_renderer.py --steps 5 > gpurun_out/r2c_bench_shape_renderer.json 2>&1; tail -c 600 gpurun_out/r2c_bench_shape_renderer.json
