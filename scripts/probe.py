"""GPU timing probe of the hot-path pieces at BASELINE cfg 2 sizes (not a bench value; prints per-call times)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glio_b200 import api, synth

W, Q, M = int(os.environ.get("W", 20)), int(os.environ.get("Q", 100000)), int(os.environ.get("M", 1000000))
t0 = time.time(); P = synth.window_problem(W=W, Q=Q, M=M); print("gen %.1fs" % (time.time() - t0))
ctx = api.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
dmap = torch.from_numpy(P["map_xyz"]).cuda(); dscans = [torch.from_numpy(s).cuda() for s in P["scans"]]
def timed(f, n=5):
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(st); f(); b.record(st); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts), ts
print("set_map ms", timed(lambda: ctx.set_map(dmap)))
ctx.window_set_scans(dscans)
nm = None
def assoc():
    global nm; nm = ctx.window_associate(P["poses_init"])
print("window_associate ms", timed(assoc)); print("matches", nm.sum(), "of", W * Q)
print("eval J ms", timed(lambda: ctx.eval_unary(P["poses_init"]), 10))
print("eval cost ms", timed(lambda: ctx.eval_unary(P["poses_init"], want_jac=False), 10))
t0 = time.time(); r = ctx.window_solve(P["poses_init"]); dt = time.time() - t0
s = r["summary"]; print("solve: iters", s.num_iterations, s.message, "wall ms %.2f" % (dt * 1e3), "cost", s.initial_cost, s.final_cost, "launches", ctx.launch_count)
print("pose err init", np.abs(P["poses_init"][:, :3] - P["poses_true"][:, :3]).max(), "final", np.abs(r["poses"][:, :3] - P["poses_true"][:, :3]).max())
