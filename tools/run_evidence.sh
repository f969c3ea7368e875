#!/bin/bash
# Round-2 evidence: every bench configuration, the config-2 volume sweep, per-layer micro-benchmarks, the ncu launch list and
# light ncu captures (batch 2 keeps ncu's per-pass memory save/restore short).  Run on the GPU box from the repo root:
#   bash tools/run_evidence.sh        (outputs under gpurun_out/, copied to profiles/ by hand)
set -u
mkdir -p gpurun_out
b() { out=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$out.err | grep '^{' | tail -1 > gpurun_out/$out.json || echo "bench $out failed"; tail -c 300 gpurun_out/$out.err; }
b r02_bench_b8 --steps 20 --warmup 5
b r02_bench_b1 --steps 20 --warmup 5 --batch 1 --no-cpu-baseline
b r02_bench_whu --steps 20 --warmup 5 --variant whu --no-cpu-baseline
b r02_bench_attonly --steps 20 --warmup 5 --att-only --batch 16 --no-cpu-baseline
b r02_bench_bf16 --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline
b r02_bench_mixed --steps 10 --warmup 3 --precision mixed --no-cpu-baseline
b r02_bench_fp32 --steps 5 --warmup 3 --precision fp32 --batch 2 --no-cpu-baseline
b r02_bench_head_b8 --steps 10 --warmup 3 --stage head --cpu-steps 1
b r02_bench_full_b8 --steps 10 --warmup 3 --stage full --cpu-steps 1
b r02_bench_reference_arm --impl reference --steps 2 --warmup 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I semstereo_b200/csrc -I include tools/probes/probe_mma_align.cu -o /tmp/probe_mma 2>/dev/null && timeout 120 /tmp/probe_mma > gpurun_out/r02_probe_mma_rate.txt
timeout 200 python tools/probes/power_probe.py 2>/dev/null | tail -5 > gpurun_out/r02_power_probe.txt
timeout 400 python tools/bench_volumes.py > gpurun_out/bench_volumes.log 2>&1; cp gpurun_out/bench_volumes.json gpurun_out/r02_bench_volumes_config2.json
timeout 200 python tools/bench_conv.py > gpurun_out/r02_bench_conv.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_split_batch2.csv python tools/ncu_path_once.py 2 split 2 > /dev/null 2>&1
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section SchedulerStats"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; timeout 400 ncu $SEC --clock-control none --import-source on -k "regex:$regex" -s $skip -c $cnt -f -o gpurun_out/$name python tools/ncu_path_once.py 2 split 2 > gpurun_out/$name.log 2>&1; tail -1 gpurun_out/$name.log; }
cap r02_ncu_conv3d 'conv3d_tc_s1f_kernel|conv3d_tc_t2_kernel|conv3d_tc_s2_kernel|conv3d_tc_s1_kernel|concat_stem_k9' 21 21
cap r02_ncu_tail 'topk_select_kernel|att_stats_kernel|ssr_upsample_kernel|sample_strength_kernel|patch_gate_blocked|regression_topk_kernel|gwc_volume|conv3d_tc_head_kernel|window_attention_core_f32|window_attn_core_mma' 11 11
ls -la gpurun_out | tail -30
