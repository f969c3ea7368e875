#!/bin/bash
# GPU job: split-route tests + regression tests + bench (default precision) with the per-kernel list
python -m pytest tests/test_gpu_split.py -x -q -s 2>&1 | tail -30
python -m pytest tests/test_gpu_tc.py tests/test_gpu_hotpath_bf16.py tests/test_gpu_surface_glue.py -x -q 2>&1 | tail -15
tag=${1:-v}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b8_split_$tag.json 2> gpurun_out/r02_b8_split_$tag.err
tail -c 600 gpurun_out/r02_b8_split_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b8_split_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']); [print(k['name'], k['ms_per_step'], k['launches_per_step'], k.get('frac')) for k in d['kernels'][:48]]"
