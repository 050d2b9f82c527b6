#!/bin/bash
# Round-2 GPU job L (8 GPUs): configs 2, 4, 5 on 8 GPUs of one box.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2l_bench_8gpu.json 2> gpurun_out/r2l_bench_8gpu.err; tail -c 500 gpurun_out/r2l_bench_8gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2l_bench_8gpu.err | tail -3
$TR scripts/bench_joint.py > gpurun_out/r2l_joint_8gpu.json 2> gpurun_out/r2l_joint.err; tail -c 700 gpurun_out/r2l_joint_8gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2l_joint.err | tail -3
$TR scripts/bench_relight.py > gpurun_out/r2l_relight_8gpu.json 2> gpurun_out/r2l_relight.err; tail -c 700 gpurun_out/r2l_relight_8gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2l_relight.err | tail -3
