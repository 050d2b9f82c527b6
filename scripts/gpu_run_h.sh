#!/bin/bash
# Round-2 GPU job H: profiles (launch lists + ncu --set full of the dominant / new kernels) and the secondary bench lines.
set -u
mkdir -p gpurun_out
python scripts/bench_material.py --steps 5 > gpurun_out/r2h_bench_material.json 2> gpurun_out/r2h_bench_material.err; tail -c 900 gpurun_out/r2h_bench_material.json
python scripts/bench_shape_renderer.py --steps 5 > gpurun_out/r2h_bench_shape_renderer.json 2>/dev/null; tail -c 400 gpurun_out/r2h_bench_shape_renderer.json
# launch lists (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2h_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2h_launches_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/r2h_launches_material.csv \
    python scripts/bench_material.py --steps 1 --warmup 3 > gpurun_out/r2h_launches_material.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file gpurun_out/r2h_launches_renderer.csv \
    python scripts/bench_shape_renderer.py --steps 1 > gpurun_out/r2h_launches_renderer.log 2>&1
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_stencil_bwd_tc -s 4 -c 1 -o gpurun_out/r2h_bwd_tc_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --check-rays 0 > gpurun_out/r2h_ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_block_bwd_tc -s 8 -c 1 -o gpurun_out/r2h_flow_bwd_full -f \
    python scripts/bench_material.py --steps 1 --warmup 3 > gpurun_out/r2h_ncu_flow_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_block_fwd_tc -s 16 -c 1 -o gpurun_out/r2h_flow_fwd_full -f \
    python scripts/bench_material.py --steps 1 --warmup 3 > gpurun_out/r2h_ncu_flow_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shader_encode_fwd -s 3 -c 1 -o gpurun_out/r2h_shader_encode_full -f \
    python scripts/bench_shape_renderer.py --steps 1 > gpurun_out/r2h_ncu_shader.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
