#!/usr/bin/env python
"""BASELINE cfg 5: kNN + point-to-plane Jacobian microbench — 1 M-point local map, 100 k query points, one pose;
K0 (grid build) + K1 (exact 5-NN, plane fit, gates) + K2 (residual + Jacobian + 6x6 reduce) as one pass.
1..N GPUs of one box (torchrun, one rank per GPU): queries sharded Q/N, map replicated (K0 runs on every GPU), one
all-reduce of the 28-double block at the end (SURVEY §8 e).  Prints one JSON line (rank 0): per-pass time (max over
ranks), achieved algorithmic GB/s per GPU (12·M + 172·Q/N bytes, SURVEY §8 d) against the measured HBM peak.
Not the driver's bench.py metric — the numbers quoted in DESIGN.md."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist_
from glio_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1000000); ap.add_argument("--Q", type=int, default=100000)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
P = synth.window_problem(W=1, Q=a.Q, M=a.M, seed=20260923 + 5)
lo, hi = rank * a.Q // world, (rank + 1) * a.Q // world
ctx = api.Context(local)
st = torch.cuda.ExternalStream(ctx.stream)
dmap = torch.from_numpy(P["map_xyz"]).cuda(); dscan = torch.from_numpy(np.ascontiguousarray(P["scans"][0][lo:hi])).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
blk = torch.zeros(28, dtype=torch.float64, device="cuda")
pose = P["poses_init"][:1]
def one_pass():
    ctx.set_map(dmap)                                   # K0
    ctx.window_set_scans([dscan]); nm = ctx.window_associate(pose)      # K1
    r = ctx.eval_unary(pose)                            # K2 (synchronises: result in pinned host memory)
    if world > 1:
        blk.copy_(torch.from_numpy(np.concatenate([r["H"].reshape(-1)[:21], r["g"].reshape(-1), r["cost"]])), non_blocking=True)
        dist_.all_reduce(blk)
    return int(nm[0])
for _ in range(a.warmup): one_pass()
ts = []
for _ in range(a.steps):
    with torch.cuda.stream(st): flush.zero_()
    torch.cuda.synchronize()
    if world > 1: dist_.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); n = one_pass(); e1.record(st); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.median(ts))], device="cuda")
if world > 1: dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
ctx.lib_profile(True)
for _ in range(5): one_pass()
ctx.synchronize(); prof = ctx.lib_profile_read(); ctx.lib_profile(False)
if rank == 0:
    peak = 6582.0
    try: peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception: pass
    byts = 12.0 * a.M + 172.0 * a.Q / world
    ms = float(t.item())
    kern = {k: round(v[0] / v[1], 4) for k, v in prof.items()}
    kms = sum(kern.values())
    print(json.dumps(dict(workload="cfg5 kNN + point-to-plane Jacobian microbench", M=a.M, Q=a.Q, n_gpus=world, queries_per_gpu=hi - lo,
                          pass_ms=ms, queries_per_s=a.Q / (ms * 1e-3), algorithmic_bytes_per_gpu=byts, achieved_GBps_per_gpu=byts / (ms * 1e-3) / 1e9,
                          frac_of_measured_hbm=byts / (ms * 1e-3) / 1e9 / peak, kernel_ms_sum=round(kms, 4), kernel_GBps=byts / (kms * 1e-3) / 1e9,
                          kernels_ms=kern, matches_rank0=n, l2="256 MB flush before every pass", timing="CUDA events on the library stream, median of %d, max over ranks" % a.steps)))
if world > 1: dist_.destroy_process_group()
