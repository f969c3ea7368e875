#!/bin/bash
# quick check after a kernel change: the tests that cover it + the default bench
python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_split.py tests/test_gpu_hotpath.py -x -q 2>&1 | grep -v Warn | tail -4
tag=${1:-q}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02_q_$tag.err | grep '^{' | tail -1 > gpurun_out/r02_q_$tag.json
tail -c 400 gpurun_out/r02_q_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/r02_q_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); [print(k['name'], k['ms_per_step'], k['launches_per_step'], k.get('frac')) for k in d['kernels'][:34]]"
