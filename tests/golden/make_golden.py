#!/usr/bin/env python
"""Generates tests/golden/window_small.npz.

WHAT THIS FIXTURE IS: a frozen input/output set of the CPU oracle (oracle/, the restatement of the reference's
algorithm) on a small seeded window problem.  The reference ships no tests or golden vectors for this path and cannot
be built here (needs ROS/PCL/Eigen/Ceres builds), so these vectors are NOT reference outputs: they pin the oracle
against drift and give the GPU box a parity target that does not need the oracle at run time.  Parity against the
reference itself stays "unpinned" (DESIGN.md section 2).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from glio_b200 import synth
from oracle import pyoracle as oracle

W, Q, M, SEED = 3, 600, 8000, 777


def problem():
    return synth.window_problem(W=W, Q=Q, M=M, seed=SEED)


def host_factor_args(P):
    """The host-side factors of the fixture problem (prior on KF 0, odometry chain, one range factor per KF)."""
    T = P["poses_true"]; rng = np.random.default_rng(SEED)
    sw = np.concatenate([np.full(3, 20.0), np.full(3, 50.0), np.full(9, 5.0)])
    out = [("prior", (0, T[0, :3] + 0.01, T[0, 3:], None, sw))]
    for i in range(W - 1):
        dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:])
        dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
        out.append(("between", (i, i + 1, dp + rng.normal(0, 0.01, 3), dq, np.zeros(3), 0.1, sw * 0.5)))
    for k in range(W):
        sat = np.array([2.0e4 * np.cos(k), 2.0e4 * np.sin(k), 2.0e4])
        out.append(("range", (k, [0.0, 0.0, 0.0], sat, float(np.linalg.norm(T[k, :3] - sat) + 0.3), 0.7)))
    return out


def oracle_outputs(P):
    prm = oracle.default_params()
    tree = oracle.KdTree(P["map_xyz"])
    res = {}
    kf, cp, nsd, score = [], [], [], []
    for k in range(W):
        t2, q2 = synth.lidar_pose_in_world(P["poses_init"][k, :3], P["poses_init"][k, 3:7])
        o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][k], t2, q2, prm=prm, tree=tree)
        v = o["status"] == oracle.GO_VALID
        res[f"status{k}"] = o["status"]; res[f"idx5_{k}"] = o["idx5"]; res[f"sqd5_{k}"] = o["sqd5"]
        res[f"nsd{k}"] = o["nsd"][v]; res[f"weight{k}"] = o["weight"][v]; res[f"src{k}"] = np.nonzero(v)[0].astype(np.int32)
        kf.append(np.full(int(v.sum()), k, np.int32)); cp.append(P["scans"][k][v]); nsd.append(o["nsd"][v]); score.append(o["score"][v])
    kf = np.concatenate(kf); cp = np.concatenate(cp); nsd = np.concatenate(nsd); score = np.concatenate(score)
    for jk in (0, 1):
        e = oracle.eval_unary(P["poses_init"], P["q_lb"], P["t_lb"], kf, cp, nsd, score, huber_delta=1.0, mode=0, jac_kind=jk)
        res[f"H_jk{jk}"] = np.stack([e["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6] for k in range(W)])
        res[f"g_jk{jk}"] = e["g"].reshape(W, 6)
        res[f"cost_jk{jk}"] = np.array([e["cost"][kf == k].sum() for k in range(W)])
    prob = oracle.WindowProblem(P["poses_init"], None, P["q_lb"], P["t_lb"], huber_delta=1.0)
    prob.add_unary(kf, cp, nsd, score)
    for kind, a in host_factor_args(P):
        getattr(prob, "add_" + kind)(*a)
    r = prob.solve(oracle.solver_options(), mode=0)
    res["solve_poses"] = r["poses"]; res["solve_steps"] = r["steps"]
    res["solve_cost"] = np.array([it["cost"] for it in r["iterations"]])
    res["solve_radius"] = np.array([it["trust_region_radius"] for it in r["iterations"]])
    res["solve_iterations"] = np.array([r["summary"].num_iterations, r["summary"].termination], np.int64)
    return res


if __name__ == "__main__":
    oracle.build()
    P = problem()
    out = dict(map_xyz=P["map_xyz"], scans=np.stack(P["scans"]), poses_init=P["poses_init"], poses_true=P["poses_true"],
               q_lb=P["q_lb"], t_lb=P["t_lb"], meta=np.array([W, Q, M, SEED], np.int64))
    out.update(oracle_outputs(P))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "window_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: int((out[f"status{k}"] == oracle.GO_VALID).sum()) for k in range(W)}, "valid;",
          "solve iterations", out["solve_iterations"].tolist())
