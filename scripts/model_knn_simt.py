#!/usr/bin/env python
"""Offline SIMT cost model of K1a (k_knn_box) on the cfg-2 workload (CPU only; numpy + scipy.spatial.cKDTree).
It replays, per warp of 32 cell-sorted queries, the two phases of the kernel in units of candidate steps:
  phase 1  3x3x3 start box: nine row ranges per lane; a warp pays max-over-lanes of each row's length
  phase 2  face growth: a lane pushes out every face nearer than its (true) 5th-neighbour distance; a warp pays, per face
           and round, the max over the lanes that need that face of the slab's candidate count
and reports the lane utilisation of each phase for alternative query orderings, to rank what to build next:
  order A  all queries of the window sorted by cell (what the kernel does today)
  order B  sorted by (coarse block of 8x8x8 cells, scan, cell): the lanes of a warp come from ONE scan, whose queries share
           their displacement from the map, so they tend to need the same faces
The model ignores the insertion cost and the exact proof margins; it is a ranking tool, not a timing prediction."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.spatial import cKDTree
from glio_b200 import synth

ap = argparse.ArgumentParser(); ap.add_argument("--W", type=int, default=20); ap.add_argument("--Q", type=int, default=100000); ap.add_argument("--M", type=int, default=1000000)
ap.add_argument("--ppc", type=float, default=8.0); ap.add_argument("--gate", type=float, default=1.5 ** 0.5)
a = ap.parse_args()
P = synth.window_problem(W=a.W, Q=a.Q, M=a.M)
m = P["map_xyz"].astype(np.float64); mn, mx = m.min(0), m.max(0); L = mx - mn + 1e-3
area = 2 * (L[0] * L[1] + L[1] * L[2] + L[0] * L[2]); cell = float(np.sqrt(a.ppc * area / len(m)))
o = mn - 0.5 * cell; dims = (np.floor((mx - o) / cell) + 2).astype(np.int64)
mc = np.floor((m - o) / cell).astype(np.int64); mid = (mc[:, 2] * dims[1] + mc[:, 1]) * dims[0] + mc[:, 0]
cnt = np.bincount(mid, minlength=int(np.prod(dims)))
cnt3 = cnt.reshape(dims[2], dims[1], dims[0])                              # [z, y, x]
csum = np.pad(cnt3, ((1, 1), (1, 1), (1, 1))).astype(np.int64)           # zero border for clipped neighbours
tree = cKDTree(m)
qs, scan_id = [], []
for k in range(a.W):
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][k, :3], P["poses_init"][k, 3:7])
    qs.append(P["scans"][k].astype(np.float64) @ synth.quat_to_R(q2).T + t2); scan_id.append(np.full(a.Q, k))
q = np.concatenate(qs); scan_id = np.concatenate(scan_id)
t0 = time.time(); r5 = tree.query(q, k=5)[0][:, 4]; print("kNN on CPU %.0f s" % (time.time() - t0), flush=True)
qc = np.floor((q - o) / cell).astype(np.int64)
inside = ((qc >= 0) & (qc < dims)).all(1)
q, qc, r5, scan_id = q[inside], qc[inside], r5[inside], scan_id[inside]
N = len(q); print("queries inside the grid:", N, "cell %.3f m" % cell)
# phase 1: nine row lengths per query (rows (dz,dy), x-range cx-1..cx+1)
x, y, z = qc[:, 0] + 1, qc[:, 1] + 1, qc[:, 2] + 1                         # +1: padded index
rows = np.stack([csum[z + dz, y + dy, x - 1] + csum[z + dz, y + dy, x] + csum[z + dz, y + dy, x + 1] for dz in (-1, 0, 1) for dy in (-1, 0, 1)], 1)
# phase 2: faces needed (true r5 vs distance to the faces of the 3x3x3 box), slab = the 9 cells beyond that face (first push only)
frac = (q - o) / cell - qc
dist_lo = (frac + 1.0) * cell; dist_hi = (2.0 - frac) * cell               # distance to the -/+ faces of the start box
r5c = np.minimum(r5, a.gate)
need = np.concatenate([dist_lo < r5c[:, None], dist_hi < r5c[:, None]], 1)   # columns: -x,-y,-z,+x,+y,+z
def slab(ax, sgn):
    idx = [z, y, x]; out = np.zeros(N, np.int64)
    for d1 in (-1, 0, 1):
        for d2 in (-1, 0, 1):
            zz, yy, xx = z.copy(), y.copy(), x.copy()
            if ax == 0: xx = np.clip(x + 2 * sgn, 0, csum.shape[2] - 1); yy = y + d1; zz = z + d2
            if ax == 1: yy = np.clip(y + 2 * sgn, 0, csum.shape[1] - 1); xx = x + d1; zz = z + d2
            if ax == 2: zz = np.clip(z + 2 * sgn, 0, csum.shape[0] - 1); xx = x + d1; yy = y + d2
            out += csum[zz, yy, xx]
    return out
slabs = np.stack([slab(0, -1), slab(1, -1), slab(2, -1), slab(0, 1), slab(1, 1), slab(2, 1)], 1)
pend = need.any(1)
print("queries needing growth: %.1f %%; faces per pending query: %.2f" % (100 * pend.mean(), need[pend].sum(1).mean()))
cid = (qc[:, 2] * dims[1] + qc[:, 1]) * dims[0] + qc[:, 0]
def evaluate(order, name):
    R = rows[order]; Nd = need[order]; Sl = slabs[order]
    nw = len(order) // 32; R = R[:nw * 32].reshape(nw, 32, 9); Nd = Nd[:nw * 32].reshape(nw, 32, 6); Sl = Sl[:nw * 32].reshape(nw, 32, 6)
    p1_warp = R.max(1).sum(); p1_lane = R.sum() / 32.0
    work = np.where(Nd, Sl, 0)
    p2_warp = work.max(1).sum(); p2_lane = work.sum() / 32.0
    print("%-34s phase1 steps/warp %6.1f (lane util %4.1f%%) | growth steps/warp %6.1f (lane util %4.1f%%) | total %6.1f" %
          (name, p1_warp / nw, 100 * p1_lane / p1_warp, p2_warp / nw, 100 * p2_lane / max(p2_warp, 1), (p1_warp + p2_warp) / nw))
evaluate(np.argsort(cid, kind="stable"), "A: cell order (today)")
blk = ((qc[:, 2] // 8) * ((dims[1] + 7) // 8) + qc[:, 1] // 8) * ((dims[0] + 7) // 8) + qc[:, 0] // 8
evaluate(np.lexsort((cid, scan_id, blk)), "B: (8^3 block, scan, cell)")
blk4 = ((qc[:, 2] // 4) * ((dims[1] + 3) // 4) + qc[:, 1] // 4) * ((dims[0] + 3) // 4) + qc[:, 0] // 4
evaluate(np.lexsort((cid, scan_id, blk4)), "B': (4^3 block, scan, cell)")
evaluate(np.lexsort((scan_id, cid)), "C: (cell, scan)")
evaluate(np.lexsort((need.dot(1 << np.arange(6)), cid)), "D: (cell, faces needed) [oracle]")
