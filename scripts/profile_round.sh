#!/bin/bash
# Every ncu capture behind profiles/r02_*: run on ONE GPU under gpurun (`gpurun --timeout 1500 -- bash scripts/profile_round.sh`),
# then summarise here with scripts/ncu_summary.py / scripts/launch_summary.py / scripts/sass_hot.py (no GPU needed for that).
# Numbers printed by anything running under ncu are never bench values.
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# launch list of the benchmark step (eager launches so that every kernel of the step is one row; shares, not absolutes, are what count)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r02_launches_step.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --skip-composite-roofline > /dev/null 2>&1
# headline compositing kernels at the roofline size (2^24 rays x 5 samples)
$NCU -k regex:composite_fwd_tile -s 1 -c 1 -o $OUT/prof_r02_composite_fwd_tile python scripts/profile_composite.py 0 shells24 > /dev/null 2>&1
$NCU -k regex:composite_bwd_tile -s 1 -c 1 -o $OUT/prof_r02_composite_bwd_tile python scripts/profile_composite.py 0 shells24 > /dev/null 2>&1
# long-ray family on BASELINE config[2]
$NCU -k regex:composite_fwd_ring -s 1 -c 1 -o $OUT/prof_r02_composite_fwd_ring python scripts/profile_composite.py 0 nerf640k > /dev/null 2>&1
$NCU -k regex:composite_bwd_ring -s 1 -c 1 -o $OUT/prof_r02_composite_bwd_ring python scripts/profile_composite.py 0 nerf640k > /dev/null 2>&1
# appearance heads, intersector, encoder at the benchmark's size
$NCU -k regex:mlp_fwd -s 2 -c 1 -o $OUT/prof_r02_mlp_fwd python scripts/profile_heads.py > /dev/null 2>&1
$NCU -k regex:mlp_bwd_stashed -s 1 -c 1 -o $OUT/prof_r02_mlp_bwd python scripts/profile_heads.py > /dev/null 2>&1
$NCU -k regex:shells_trace -s 3 -c 1 -o $OUT/prof_r02_shells_trace python scripts/bench_trace.py > /dev/null 2>&1
# encoder kernels on the benchmark's real hit points (packed order; the backward with key = samples_layer as the step calls it)
$NCU -k regex:permuto_fwd -s 3 -c 1 -o $OUT/prof_r02_permuto_fwd python scripts/bench_permuto_hits.py > /dev/null 2>&1
$NCU -k regex:permuto_bwd -s 45 -c 1 -o $OUT/prof_r02_permuto_bwd python scripts/bench_permuto_hits.py > /dev/null 2>&1
ls -la $OUT/*.ncu-rep $OUT/r02_launches_step.csv
