#!/usr/bin/env python
"""Batch (scan-to-multiscan) path at BASELINE cfg 3 / cfg 4 shape, 1..N GPUs of one box (torchrun, one rank per GPU).
Keyframe-sharded: every rank declares the same pair list, holds its keyframe range + halo, associates and evaluates
only its own pairs; the pose-block buffers are summed with one NCCL all-reduce per buffer per evaluation; the banded
solve is replicated.  Prints one JSON line (rank 0).  Not the driver's bench.py metric — the numbers quoted in DESIGN.md."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist_
from glio_b200 import api, dist, synth

ap = argparse.ArgumentParser()
ap.add_argument("--K", type=int, default=200); ap.add_argument("--Q", type=int, default=100000)
ap.add_argument("--search-range", type=int, default=6); ap.add_argument("--sel", type=int, default=0, help="keep the first n matches per pair (0 = no selection)")
ap.add_argument("--solves", type=int, default=3); ap.add_argument("--max-iter", type=int, default=100)
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
n_bound = None
if os.environ.get("GLIO_BIND", "1") != "0":             # host threads on the CPUs local to this rank's GPU (as bench.py does)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    n_bound = bench.bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
if world > 1:
    dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
K, sr = a.K, a.search_range
cur, oth = dist.batch_pairs(K, sr)
own = dist.owner_of(cur, K, world) == rank
need = dist.frames_needed(cur, oth, own)
t0 = time.time(); B = synth.batch_problem(K=K, Q=a.Q, search_range=sr, frames=need); t_gen = time.time() - t0
ctx = api.Context(local)
hook = None
if world > 1:
    hook = dist.NcclHook(rank, world); hook.install(ctx)
st = torch.cuda.ExternalStream(ctx.stream)
def sync_all():
    torch.cuda.synchronize()
    if world > 1: dist_.barrier()
# upload frames (resident), declare the global pair list
dscans = {int(k): torch.from_numpy(B["scans"][k]).cuda() for k in need}
for k in need: ctx.batch_set_frame(int(k), dscans[int(k)], B["poses_init"][k])
ctx.batch_declare_pairs(cur, oth)
def associate():
    return ctx.batch_associate_pairs(cur[own], oth[own])
associate(); sync_all()                                   # warm-up (allocations)
for k in need: ctx.batch_set_pose(int(k), B["poses_init"][k])   # invalidate the grids so the timed pass rebuilds them
sync_all(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st); nm = associate(); e1.record(st); sync_all()
t_assoc = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1: dist_.all_reduce(t_assoc, op=dist_.ReduceOp.MAX)
if a.sel > 0:
    for (c, o), n in zip(zip(cur[own], oth[own]), nm): ctx.batch_select(int(c), int(o), np.arange(min(a.sel, n), dtype=np.int32))
nres = torch.tensor([float(np.minimum(nm, a.sel).sum() if a.sel > 0 else nm.sum())], device="cuda", dtype=torch.float64)
if world > 1: dist_.all_reduce(nres)
# host factors: prior on KF 0 + odometry-like chain (stand-ins for the IMU chain / delta_q factors of the batch problem)
hf = api.HostFactorSet(); T = B["poses_true"]; rng = np.random.default_rng(7)
sw = np.concatenate([np.full(3, 10.0), np.full(3, 30.0), np.zeros(9)])
hf.add_prior(0, T[0, :3], T[0, 3:], None, sw * 3)
for i in range(K - 1):
    dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:]); dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
    hf.add_between(i, i + 1, dp + rng.normal(0, 0.005, 3), dq, np.zeros(3), 0.1, sw)
opt = api.batch_solver_options(max_num_iterations=a.max_iter)
r = ctx.batch_solve(B["poses_init"], None, hf, opt)         # warm-up
sync_all(); e0.record(st); t0 = time.perf_counter(); iters = 0
for _ in range(a.solves):
    r = ctx.batch_solve(B["poses_init"], None, hf, opt); iters += len(r["steps"])
e1.record(st); sync_all(); wall = time.perf_counter() - t0
t_solve = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1: dist_.all_reduce(t_solve, op=dist_.ReduceOp.MAX)
# all-reduce latency of the two block buffers alone
ar_us = None
if world > 1:
    P = len(cur); b1 = torch.zeros(K * 28, dtype=torch.float64, device="cuda"); b2 = torch.zeros(P * 36, dtype=torch.float64, device="cuda")
    for _ in range(5): dist_.all_reduce(b1); dist_.all_reduce(b2)
    torch.cuda.synchronize(); f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(50): dist_.all_reduce(b1); dist_.all_reduce(b2)
    f1.record(); torch.cuda.synchronize(); ar_us = f0.elapsed_time(f1) / 50 * 1e3
if rank == 0:
    s = r["summary"]
    print(json.dumps(dict(workload="batch sms_fusion_level=1 scan-to-multiscan", K=K, Q=a.Q, search_range=sr, n_gpus=world, pairs=int(len(cur)),
                          queries=int(len(cur)) * a.Q, residuals=float(nres.item()), selection=a.sel, assoc_ms=float(t_assoc.item()),
                          assoc_Mqueries_per_s=len(cur) * a.Q / (t_assoc.item() * 1e-3) / 1e6, iterations=iters / a.solves, evaluations=s.num_evaluations,
                          solve_ms=float(t_solve.item()) / a.solves, iterations_per_s=iters / (t_solve.item() * 1e-3), host_wall_ms=1e3 * wall / a.solves,
                          allreduce_us_per_eval=ar_us, termination=s.message.decode(), cost=[s.initial_cost, s.final_cost], gen_s=round(t_gen, 1),
                          pose_err=[float(np.abs(B["poses_init"][:, :3] - T[:, :3]).max()), float(np.abs(r["poses"][:, :3] - T[:, :3]).max())])))
# host share of one solve: kernel time (CUDA events per launch) vs the solver's own wall-clock timers
ctx.lib_profile(True)
t0 = time.perf_counter(); r = ctx.batch_solve(B["poses_init"], None, hf, opt); wall1 = time.perf_counter() - t0
prof = ctx.lib_profile_read(); ctx.lib_profile(False)
if rank == 0:
    s = r["summary"]
    print(json.dumps(dict(host_share=dict(cpus_bound=n_bound, python_wall_ms=1e3 * wall1, solver_total_ms=1e3 * s.total_seconds, eval_ms=1e3 * s.eval_seconds,
                                          linear_solver_ms=1e3 * s.linear_solver_seconds, evaluations=s.num_evaluations, jacobian_evaluations=s.num_jacobian_evaluations,
                                          linear_solves=s.num_linear_solves, kernels_ms={k: round(v[0], 3) for k, v in prof.items()},
                                          kernel_launches={k: v[1] for k, v in prof.items()}))))
if hook: hook.close()
if world > 1: dist_.destroy_process_group()
