#!/bin/bash
# Runs on the GPU box (under gpurun): the evidence the roofline numbers in bench.py come from.
#   1. --set full captures of the dominant kernels (k_knn_box, k_plane_fit, k_eval_unary) -> profiles/ncu_traffic.json,
#      which bench.py reads for roofline.traffic (DRAM bytes per launch) and the issue-bound evidence
#   2. the bench line itself and the reference arm (never measured under a profiler)
#   3. per-launch device times of one bench command (ncu, cold-cache & serialised: compare SHARES)
# Afterwards, here: python scripts/summarize_profiles.py <round>   (gpurun_out/ -> profiles/)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_knn_box|k_plane_fit|k_eval_unary" -c 16 -f -o gpurun_out/prof_top python scripts/probe.py > gpurun_out/ncu_top.log 2>&1
python scripts/summarize_profiles.py > gpurun_out/summarize.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_line.json 2>> gpurun_out/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
ls -la gpurun_out
