import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from glio_b200 import api, synth
P = synth.window_problem(W=20, Q=100000, M=1000000)
ctx = api.Context(0)
dmap = torch.from_numpy(P["map_xyz"]).cuda(); dscans = [torch.from_numpy(s).cuda() for s in P["scans"]]
for i in range(4):
    ctx.set_map(dmap); ctx.window_set_scans(dscans)
    if i == 3: os.environ["GLIO_TRACE"] = "1"
    t0 = time.perf_counter(); ctx.window_associate(P["poses_init"]); print("associate wall ms", 1e3 * (time.perf_counter() - t0))
os.environ.pop("GLIO_TRACE")
hf = api.HostFactorSet()
sb0 = np.zeros((20, 9))
import cProfile, pstats
r = ctx.window_solve(P["poses_init"], sb0, hf, band=29)
t0 = time.perf_counter()
for _ in range(20): r = ctx.window_solve(P["poses_init"], sb0, hf, band=29)
print("solve wall ms", 1e3 * (time.perf_counter() - t0) / 20, "iters", r["summary"].num_iterations, "evals", r["summary"].num_evaluations)
t0 = time.perf_counter()
for _ in range(50): ctx.eval_unary(P["poses_init"])
print("eval_unary wall ms", 1e3 * (time.perf_counter() - t0) / 50)
