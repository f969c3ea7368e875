#!/bin/bash
# Round evidence: tests, smoke, every bench configuration, per-layer micro-benchmarks, ncu launch list + full captures.
# Run on the GPU box from the repo root:  bash tools/run_evidence.sh   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep "\[smoke\]\|Error\|assert"
b() { out=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err || echo "bench $out failed"; tail -c 200 gpurun_out/$out.err; }
b bench_b8 --steps 20 --warmup 5
b bench_b1 --steps 20 --warmup 5 --batch 1 --no-cpu-baseline
b bench_whu --steps 20 --warmup 5 --variant whu --no-cpu-baseline
b bench_attonly --steps 20 --warmup 5 --att-only --batch 16 --no-cpu-baseline
b bench_fp32 --steps 5 --warmup 3 --precision fp32 --batch 2 --no-cpu-baseline
b bench_head --steps 10 --warmup 3 --stage head --cpu-steps 1
b bench_ref --impl reference --steps 2 --warmup 1
timeout 200 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1
timeout 200 python tools/bench_conv2d.py 8 > gpurun_out/bench_conv2d.log 2>&1
timeout 200 python tools/bench_decoder.py 8 > gpurun_out/bench_decoder.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline --no-graph > /dev/null 2>&1
cap() { name=$1; regex=$2; skip=$3; shift 3; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip --launch-count 1 -o gpurun_out/$name "$@" > /dev/null 2>&1; }
cap k9_concat_stem concat_stem_k9_kernel 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph
cap s1f_classif0 conv3d_tc_s1f_kernel 10 python tools/bench_conv.py "classif.0[folded]"
cap head_classif2 conv3d_tc_head_kernel 3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph
cap attn_core window_attn_core_mma_kernel 3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph
cap sample_strength sample_strength_kernel 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph
ls -la gpurun_out | tail -30
