#!/bin/bash
# Round 2, final kernels: launch list + --set full captures of the dominant kernels (run under gpurun; outputs in gpurun_out/,
# summarised by scripts/summarize_profiles.py 2).  Never a bench value.
set -x
B="python bench.py --no-cpu-baseline --no-batch --no-microbench"
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv $B --steps 4 --warmup 3 > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_knn_box|k_plane_fit|k_eval_unary" -s 18 -c 6 -o gpurun_out/prof_top $B --steps 2 --warmup 3 > gpurun_out/ncu_top.log 2>&1
SWEEP="8:2:0,8:2:1,8:4,8:5,8:6" python scripts/sweep_knn.py > gpurun_out/sweep_modes.log 2>&1
tail -2 gpurun_out/ncu_top.log; cat gpurun_out/sweep_modes.log
