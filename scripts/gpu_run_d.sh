#!/bin/bash
# Round-2 GPU job D: flow / material / prefilter tests after the fused coupling-block kernels, material bench.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_flow_gpu.py tests/test_golden.py tests/test_mc_gpu.py tests/test_prefilter_gpu.py tests/test_nvs_gpu.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2d_pytest.log
tail -40 gpurun_out/r2d_pytest.log
python scripts/bench_material.py --steps 5 > gpurun_out/r2d_bench_material.json 2> gpurun_out/r2d_bench_material.err; tail -c 1800 gpurun_out/r2d_bench_material.json; tail -3 gpurun_out/r2d_bench_material.err
