#!/bin/bash
python -m pytest tests/test_gpu_decoder.py tests/test_gpu_surface_glue.py tests/test_gpu_ops.py tests/test_gpu_split.py -x -q -s 2>&1 | grep -v Warning | tail -30
