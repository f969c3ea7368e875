#!/bin/bash
python -m pytest tests/test_gpu_backbone.py tests/test_gpu_surface_glue.py tests/test_gpu_pipeline.py -x -q -s 2>&1 | tail -30
python bench.py --steps 10 --warmup 3 --stage full --cpu-steps 1 > gpurun_out/r02_full_b8.json 2> gpurun_out/r02_full_b8.err
tail -c 800 gpurun_out/r02_full_b8.err
python -c "
import json; d=json.load(open('gpurun_out/r02_full_b8.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline')); [print(k['name'], k['ms_per_step'], k['launches_per_step'], k.get('frac')) for k in d['kernels'][:30]]"
