#!/bin/bash
python -m pytest tests/test_gpu_train.py -x -q 2>&1 | grep -v Warning | tail -25
python tools/train_step.py --steps 3 --warmup 2 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
tail -c 1500 gpurun_out/r02_train_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_train_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['library_ms_per_step'], d['final_loss']); [print(k) for k in d['kernels'][:14]]"
