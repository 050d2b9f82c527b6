#!/bin/bash
# Round-2 GPU job J (2 GPUs): shader / renderer tests after the coalesced encode stores, then the data-parallel paths.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_shader.py tests/test_renderer.py -q -m gpu 2>&1 | tail -4
python scripts/bench_shape_renderer.py --steps 10 > gpurun_out/r2j_bench_shape_renderer.json 2>/dev/null; tail -c 700 gpurun_out/r2j_bench_shape_renderer.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; tail -c 900 gpurun_out/r2j_bench_2gpu.json; tail -3 gpurun_out/r2j_bench_2gpu.err
$TR scripts/bench_joint.py --check --rays 2048 --samples 64 --grid 64 --mat-grid 64 --tris-u 100 --tris-v 51 --micro 512 > gpurun_out/r2j_joint_check_2gpu.json 2> gpurun_out/r2j_joint_check.err; tail -c 600 gpurun_out/r2j_joint_check_2gpu.json; tail -3 gpurun_out/r2j_joint_check.err
$TR scripts/bench_joint.py > gpurun_out/r2j_joint_2gpu.json 2> gpurun_out/r2j_joint.err; tail -c 600 gpurun_out/r2j_joint_2gpu.json; tail -3 gpurun_out/r2j_joint.err
