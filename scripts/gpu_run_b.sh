#!/bin/bash
# Round-2 GPU job B: full GPU suite (no -x: collect every failure), prefilter goldens from the reference kernels, tcgen05
# probe table, short bench.
set -u
mkdir -p gpurun_out
python oracle/gen_golden_prefilter.py gpurun_out/prefilter.npz > gpurun_out/r2b_prefilter.log 2>&1; tail -2 gpurun_out/r2b_prefilter.log
cp gpurun_out/prefilter.npz tests/golden/prefilter.npz 2>/dev/null
python -m pytest tests/ -q -m gpu 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
timeout 900 python scripts/tc_rate_probe.py gpurun_out/r2b_tc_probe.json > gpurun_out/r2b_tc_probe.log 2>&1
python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/r2b_bench.err | tee gpurun_out/r2b_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'])
print('check', d['check'])
print('secondary', d['secondary'])
print(d['roofline']['kernels_ms_per_step'])"
tail -5 gpurun_out/r2b_bench.err
