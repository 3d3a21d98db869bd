#!/usr/bin/env python
"""Local map maintenance on the device at the reference's sizes: local_map_width = 50 keyframe clouds of Q
surf points resident on the GPU; per new keyframe: pop_front + push (one cloud crosses PCIe) + build (concatenate,
VoxelGrid 0.4 m, grid build).  Timed beside the host-side alternative it replaces on the GPU path: upload the finished
down-sampled map and glio_set_map.  Prints one JSON line.  Not the driver's bench.py metric."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from glio_b200 import api, synth

ap = argparse.ArgumentParser(); ap.add_argument("--K", type=int, default=50); ap.add_argument("--Q", type=int, default=100000)
ap.add_argument("--leaf", type=float, default=0.4); ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
rng = np.random.default_rng(5); sc = synth.Scene(-60.0, a.K + 60.0, rng); poses = synth.trajectory(a.K + a.reps + 2, rng)
clouds = [synth.scan_in_lidar_frame(sc, poses[k], a.Q, rng) for k in range(a.K + a.reps + 2)]
pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
ctx = api.Context(0); st = torch.cuda.ExternalStream(ctx.stream)
def T(k): return synth.lidar_pose_in_world(poses[k, :3], poses[k, 3:7])
for k in range(a.K): ctx.localmap_push(pinned[k].numpy(), *T(k))
n_map = ctx.localmap_build(a.leaf); ctx.localmap_build(a.leaf)
ts = []
for r in range(a.reps):
    k = a.K + r
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); ctx.localmap_pop_front(); ctx.localmap_push(pinned[k].numpy(), *T(k)); n_map = ctx.localmap_build(a.leaf); e1.record(st)
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ctx.lib_profile(True)
for r in range(3): ctx.localmap_build(a.leaf)
ctx.synchronize(); prof = {k: round(v[0] / v[1], 4) for k, v in ctx.lib_profile_read().items()}; ctx.lib_profile(False)
m = ctx.get_map(); pm = torch.from_numpy(m).pin_memory().numpy()
th = []
for r in range(a.reps):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); ctx.set_map(pm); e1.record(st); torch.cuda.synchronize(); th.append(e0.elapsed_time(e1))
nf, npts = ctx.localmap_size()
# CPU beside it: the oracle's restatement of pcl::VoxelGrid (literal std::sort variant, 1 thread as PCL) on the same raw cloud
from oracle import pyoracle as po
po.build()
world = np.concatenate([po.transform_points(clouds[k], *T(k)) for k in range(a.reps, a.K + a.reps)])
t0 = time.perf_counter(); ref, _ = po.voxel_filter(world, a.leaf, stable=False); cpu_ms = 1e3 * (time.perf_counter() - t0)
print(json.dumps(dict(workload="local map maintenance on device", frames=nf, raw_points=npts, leaf=a.leaf, map_points=n_map,
                      device_update_ms=float(np.median(ts)), device_update_what="pop_front + push one %d-point cloud from pinned host memory + concatenate + voxel filter + grid build" % a.Q,
                      host_map_upload_plus_set_map_ms=float(np.median(th)), cpu_voxel_filter_ms=cpu_ms, cpu_what="oracle port of pcl::VoxelGrid on the same %d raw points, 1 thread (concatenation and kd-tree build not included)" % len(world), cpu_map_points=int(len(ref)), h2d_bytes_device_path=12 * a.Q, h2d_bytes_host_path=12 * int(n_map), kernels_ms=prof)))
