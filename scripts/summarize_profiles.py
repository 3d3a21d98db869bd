#!/usr/bin/env python
"""Turns the raw captures of scripts/collect_profiles.sh (gpurun_out/) into the tracked summaries under profiles/:
  profiles/rNN_launch_list.csv / _summary.txt   per-launch device times of one bench command under ncu (shares only)
  profiles/rNN_ncu_top_kernels.txt              selected --set full metrics of the dominant kernels
  profiles/rNN_bench_line.json, rNN_bench_reference_line.json   the bench lines of the same box (not under a profiler)
Run here after the gpurun call:  python scripts/summarize_profiles.py [round]"""
import csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
rnd = "r%02d" % int(sys.argv[1]) if len(sys.argv) > 1 else "r01"

# ---- launch list ----
src = os.path.join(G, "launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, rnd + "_launch_list.csv"))
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    h = rows[0]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = re.sub(r"^(void )?(glio::)?", "", r[kn]); name = re.sub(r"\(.*$", "", name)[:48]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, rnd + "_launch_list_summary.txt"), "w") as f:
        f.write("# round %s — per-launch device time of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under\n" % rnd[1:])
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400   (cold-cache, serialised launches: compare SHARES)\n")
        f.write("# (first 400 launches: warm-up steps + the start of the timed region)\n")
        f.write("%-48s %9s %12s %8s\n" % ("kernel", "launches", "total_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-48s %9d %12.1f %7.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
    print("wrote launch list summary (%d launches)" % sum(v[0] for v in agg.values()))

# ---- --set full captures ----
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
rep = os.path.join(G, "prof_top.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); h, units = rows[0], rows[1]
    with open(os.path.join(P, rnd + "_ncu_top_kernels.txt"), "w") as f:
        f.write("# round %s — ncu --set full --clock-control none --import-source on (one capture per kernel, B200, cfg 2 sizes: M=1M map, 2M queries, 2.0M residuals)\n" % rnd[1:])
        last = {}
        for r in rows[2:]: last[r[h.index("Kernel Name")]] = r            # several launches are captured: keep the last (warm) one of each kernel
        for r in last.values():
            f.write("\n## %s\n" % r[h.index("Kernel Name")][:90])
            for w in WANT:
                if w in h: f.write("  %-100s %16s %s\n" % (w, r[h.index(w)], units[h.index(w)]))
    print("wrote ncu summary (%d kernels of %d captured launches)" % (len(last), len(rows) - 2))

# ---- DRAM traffic per launch of the roofline kernels (bench.py reads profiles/ncu_traffic.json into roofline.traffic) ----
def _traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3: return {}
    h, u = rows[0], rows[1]; res = {}
    def num(r, m):
        return float(r[h.index(m)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[h.index(m)], 1)
    for r in rows[2:]:
        name = re.sub(r"^(void )?(glio::)?", "", r[h.index("Kernel Name")]); name = re.sub(r"[<(].*$", "", name)
        e = res.setdefault(name, dict(dram_bytes=[], warp_inst=[], issue_active_pct=[], threads_per_inst=[], time_us=[]))
        e["dram_bytes"].append(num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"))
        e["warp_inst"].append(num(r, "smsp__inst_executed.sum")); e["issue_active_pct"].append(num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"))
        e["threads_per_inst"].append(num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")); e["time_us"].append(num(r, "gpu__time_duration.sum"))
    return {k: {m: round(sum(v) / len(v), 2) for m, v in e.items()} for k, e in res.items()}
tr = {}
for rep in ("prof_top.ncu-rep",):
    if os.path.exists(os.path.join(G, rep)): tr.update(_traffic(os.path.join(G, rep)))
if tr:
    tr["_source"] = "ncu --set full --clock-control none per launch, cfg 2 sizes (gpurun_out/prof_top.ncu-rep of round %s, mean over the captured launches); dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum" % rnd[1:]
    json.dump(tr, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1); print("wrote ncu_traffic.json", {k: v for k, v in tr.items() if k != "_source"})

for name in ("bench_line.json", "bench_reference_line.json"):
    s = os.path.join(G, name)
    if os.path.exists(s) and os.path.getsize(s) > 10:
        json.loads(open(s).read().strip().splitlines()[-1])           # must be one valid JSON line
        shutil.copy(s, os.path.join(P, rnd + "_" + name)); print("copied", name)
