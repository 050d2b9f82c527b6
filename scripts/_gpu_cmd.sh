python -m pytest tests/test_occ_gpu.py tests/test_shape_gpu.py tests/test_renderer.py tests/test_optim_gpu.py tests/test_nvs_gpu.py -x -q -m gpu 2>&1 | tail -6
python scripts/bench_adam.py | tee gpurun_out/bench_adam_r1.json
python bench.py 2> gpurun_out/bench_r1i.err | tee gpurun_out/bench_r1i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['calls_ms'], d['roofline']['kernels_ms_per_step'])"
