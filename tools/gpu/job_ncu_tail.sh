#!/bin/bash
# ncu captures of the memory-bound tail kernels + the split-route helpers, one launch each (second pass of the path, batch 2 to keep
# ncu's per-pass memory save/restore cheap)
mkdir -p gpurun_out
K='topk_select_kernel|att_stats_kernel|ssr_upsample_kernel|sample_strength_kernel|patch_gate_blocked|regression_topk_kernel|gwc_volume|conv3d_tc_head_kernel|window_attention_core_f32'
B=${1:-2}
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats \
  --clock-control none --import-source on -k "regex:$K" -s 10 -c 10 -f -o gpurun_out/r02_tail python tools/ncu_path_once.py $B split 2 > gpurun_out/r02_tail.log 2>&1
tail -3 gpurun_out/r02_tail.log
ncu -i gpurun_out/r02_tail.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
print(' | '.join(keys))
for r in rows[2:]:
    d=dict(zip(h,r)); print(' | '.join(str(d.get(k,''))[:48] for k in keys))
"
