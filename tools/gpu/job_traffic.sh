#!/bin/bash
# one `ncu --set full` capture of the dominant kernel (concat_stem with the volume generated in-kernel) at batch 8: DRAM traffic per launch
timeout 800 ncu --set full --clock-control none --import-source on -k regex:concat_stem_k9_kernel -s 1 -c 1 -f -o gpurun_out/r02_ncu_k9_concat_stem_b8 python tools/ncu_path_once.py 8 split 2 > gpurun_out/r02_ncu_k9.log 2>&1
tail -2 gpurun_out/r02_ncu_k9.log
python tools/ncu_summary.py gpurun_out/r02_ncu_k9_concat_stem_b8.ncu-rep | head -40
