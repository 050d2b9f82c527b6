#!/bin/bash
# Round-2 GPU job A: full GPU test suite (+ achieved-error record), tcgen05 layout / rate probes, sanitizer passes on the
# scatter / scan kernels, bench with parity check + secondary lines.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log | tail -5
timeout 600 python scripts/tc_rate_probe.py gpurun_out/r2a_tc_probe.json > gpurun_out/r2a_tc_probe.log 2>&1
tail -3 gpurun_out/r2a_tc_probe.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py -q -x \
    -k "vm_feature or neus_composite or tv_loss or gaussian" > gpurun_out/r2a_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2a_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_shape_gpu.py -q -x \
    -k "neus_composite or tv_loss or gaussian" > gpurun_out/r2a_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2a_sanitizer_racecheck.log
python bench.py 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'])
print('check', d['check'])
print('secondary', d['secondary'])
print(d['roofline']['kernels_ms_per_step'])"
tail -5 gpurun_out/r2a_bench.err
