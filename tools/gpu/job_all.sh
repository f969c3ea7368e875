#!/bin/bash
# GPU job: the whole GPU suite + smoke + bench (default precision)
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
tag=${1:-v}
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_b8_split_$tag.json 2> gpurun_out/r02_b8_split_$tag.err
tail -c 600 gpurun_out/r02_b8_split_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/r02_b8_split_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline')); [print(k['name'], k['ms_per_step'], k['launches_per_step'], k.get('frac')) for k in d['kernels'][:20]]"
