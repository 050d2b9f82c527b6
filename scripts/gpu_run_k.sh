#!/bin/bash
# Round-2 GPU job K (2 GPUs): new material tests, joint step on 2 GPUs (gradient check + timing).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_mc_gpu.py -q -m gpu -k "million or bench_scale or train_step" 2>&1 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR scripts/bench_joint.py --check --rays 2048 --samples 64 --grid 64 --mat-grid 64 --tris-u 100 --tris-v 51 --micro 512 > gpurun_out/r2k_joint_check_2gpu.json 2> gpurun_out/r2k_joint_check.err; tail -c 600 gpurun_out/r2k_joint_check_2gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2k_joint_check.err | tail -5
$TR scripts/bench_joint.py > gpurun_out/r2k_joint_2gpu.json 2> gpurun_out/r2k_joint.err; tail -c 600 gpurun_out/r2k_joint_2gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2k_joint.err | tail -5
