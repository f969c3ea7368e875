#!/bin/bash
# ncu --set full of the three memory-side tail kernels at batch 8 (second pass of the path), details page printed to a text file
mkdir -p gpurun_out
K=${1:-'sample_strength_kernel|ssr_upsample_kernel|topk_select_kernel'}
N=${2:-3}
tag=${3:-three}
timeout 800 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s $N -c $N -f -o gpurun_out/r02_$tag python tools/ncu_path_once.py 8 split 2 > gpurun_out/r02_$tag.log 2>&1
tail -2 gpurun_out/r02_$tag.log
ncu -i gpurun_out/r02_$tag.ncu-rep --page details > gpurun_out/r02_${tag}_details.txt 2>/dev/null
wc -l gpurun_out/r02_${tag}_details.txt
