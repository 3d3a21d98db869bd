"""K3: marginalisation of the oldest keyframe (MarginalizationInfo, GLIO/src/MarginalizationFactor.cpp:82-202; call site
Estimator.cpp:2462-2608).
CPU: the Schur / eigen step against numpy; the product's host factors (marginalisation path: ambient x,y,z quaternion columns)
against the oracle's Jet evaluation; the product's MargPrior factor in the solve path against the oracle's literal
MarginalizationFactor::Evaluate.
GPU: the integrated pass (device LiDAR blocks + host factors -> Schur -> prior) against the oracle's literal restatement, two
windows in a row so that the prior of window 1 is consumed by the solve AND by the marginalisation of window 2."""
import ctypes

import numpy as np
import pytest

from glio_b200 import synth


def test_marginalize_matches_numpy_invariants():
    from glio_b200 import api
    lib = api.lib()
    rng = np.random.default_rng(0)
    N, m = 138, 15                       # 6W+18 at W = 20, drop KF0's (t, q, speed-bias)
    J = rng.normal(size=(400, N)); J[:, 20:40] *= 1e-3
    J[:, 100:103] = 0.0                  # unobservable directions -> eigenvalues below eps are truncated
    A = J.T @ J; b = J.T @ rng.normal(size=400)
    n = N - m
    LJ = np.zeros((n, n)); lr = np.zeros(n)
    rc = lib.glio_marginalize(A.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(N), ctypes.c_int(m),
                              ctypes.c_double(1e-8), LJ.ctypes.data_as(ctypes.c_void_p), lr.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    # numpy restatement of MarginalizationFactor.cpp:176-201
    eps = 1e-8
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    Ainv = V @ np.diag(np.where(w > eps, 1 / np.where(w > eps, w, 1), 0)) @ V.T
    Ar = A[m:, m:] - A[m:, :m] @ Ainv @ A[:m, m:]
    br = b[m:] - A[m:, :m] @ Ainv @ b[:m]
    w2, V2 = np.linalg.eigh(Ar)
    S = np.where(w2 > eps, w2, 0); Si = np.where(w2 > eps, 1 / np.where(w2 > eps, w2, 1), 0)
    Jref = np.diag(np.sqrt(S)) @ V2.T; rref = np.diag(np.sqrt(Si)) @ V2.T @ br
    assert (w2 <= eps).sum() >= 3        # the truncation branch is exercised
    scale = np.abs(Ar).max()
    assert np.max(np.abs(LJ.T @ LJ - Jref.T @ Jref)) <= 1e-9 * scale          # J^T J (invariant to eigenvector signs/order)
    assert np.max(np.abs(LJ.T @ lr - Jref.T @ rref)) <= 1e-9 * np.abs(Jref.T @ rref).max()
    assert abs(lr @ lr - rref @ rref) <= 1e-9 * (rref @ rref)


def test_marginalize_block_diagonal_components_equal_dense():
    """The component-wise eigen-decomposition (what makes the window's block-diagonal Schur complement cheap) gives the same
    prior as numpy's dense one, and keeps exact zeros between uncoupled keyframes."""
    from glio_b200 import api
    lib = api.lib()
    rng = np.random.default_rng(3)
    W = 8; N = 6 * W + 18; m = 15; n = N - m
    A = np.zeros((N, N)); b = rng.normal(size=N)
    def add(idx, rows):
        Jb = rng.normal(size=(rows, len(idx))); A[np.ix_(idx, idx)] += Jb.T @ Jb
    add(list(range(0, 30)), 40)                       # KF0 <-> KF1 (IMU-like)
    for k in range(W):                                 # unary LiDAR blocks
        base = 0 if k == 0 else (15 if k == 1 else 30 + 6 * (k - 2))
        add(list(range(base, base + 6)), 30)
    LJ = np.zeros((n, n)); lr = np.zeros(n)
    assert lib.glio_marginalize(A.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(N), ctypes.c_int(m), ctypes.c_double(1e-8),
                                LJ.ctypes.data_as(ctypes.c_void_p), lr.ctypes.data_as(ctypes.c_void_p)) == 0
    Ainv = np.linalg.inv(0.5 * (A[:m, :m] + A[:m, :m].T))
    Ar = A[m:, m:] - A[m:, :m] @ Ainv @ A[:m, m:]; br = b[m:] - A[m:, :m] @ Ainv @ b[:m]
    assert np.max(np.abs(LJ.T @ LJ - Ar)) <= 1e-10 * np.abs(Ar).max()
    assert np.max(np.abs(LJ.T @ lr - br)) <= 1e-10 * np.abs(br).max()
    G = LJ.T @ LJ
    assert np.all(G[:15, 15:] == 0.0) and np.all(G[15:21, 21:] == 0.0)        # exact zeros survive: the band of the next solve stays narrow


def _window_with_factors(oracle, W=5, Q=1200, M=30000, seed=91):
    P = synth.window_problem(W=W, Q=Q, M=M, seed=seed)
    rng = np.random.default_rng(seed)
    T = P["poses_true"]
    sb = rng.normal(0, 0.05, (W, 9))
    sw = np.concatenate([np.full(3, 20.0), np.full(3, 50.0), np.full(9, 5.0)])
    spec = dict(prior=(0, T[0, :3] + 0.01, T[0, 3:], rng.normal(0, 0.02, 9), sw), between=[])
    for i in range(W - 1):
        dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:])
        dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
        spec["between"].append((i, i + 1, dp + rng.normal(0, 0.01, 3), dq, rng.normal(0, 0.02, 3), 0.1, sw * 0.5))
    return P, sb, spec


def _oracle_problem(oracle, P, sb, spec, poses, matches):
    prob = oracle.WindowProblem(poses, sb, P["q_lb"], P["t_lb"], huber_delta=1.0)
    for k, (cp, nsd, score) in enumerate(matches):
        prob.add_unary(np.full(len(cp), k, np.int32), cp, nsd, score)
    prob.add_prior(*spec["prior"])
    for bfac in spec["between"]:
        prob.add_between(*bfac)
    return prob


def _product_factors(spec):
    from glio_b200 import api
    hf = api.HostFactorSet()
    hf.add_prior(*spec["prior"])
    for bfac in spec["between"]:
        hf.add_between(*bfac)
    return hf


def test_host_factor_marg_path_matches_oracle_jets(oracle):
    """glio_hf_marg_evaluate (analytic ambient x,y,z columns of the stand-in prior / between factors) against the oracle's
    dual-number evaluation pushed through ThreadsConstructA's rightCols(3) rule."""
    P, sb, spec = _window_with_factors(oracle)
    poses = P["poses_init"]
    hf = _product_factors(spec)
    A, b = hf.marg_evaluate(poses, sb)
    prob = _oracle_problem(oracle, P, sb, spec, poses, [])
    o = prob.marginalize()
    assert np.max(np.abs(A - o["A"])) <= 1e-11 * np.abs(o["A"]).max()
    assert np.max(np.abs(b - o["b"])) <= 1e-11 * np.abs(o["b"]).max()
    assert np.abs(A[:30, :30]).max() > 0 and np.all(A[30:, :] == 0)        # only KF0 / KF1 are touched by these factors


def test_marg_prior_factor_solve_path_matches_oracle(oracle):
    """A MargPrior inside the host factor set (information form) against the oracle's literal MarginalizationFactor::Evaluate
    (residual = r0 + J dx, analytic quaternion Jacobian, Ceres' plus-Jacobian): same cost, J^T J and J^T r."""
    from glio_b200 import api
    rng = np.random.default_rng(5)
    W = 4; n = 6 * W + 3
    LJ = rng.normal(size=(n, n)) * (rng.random((n, n)) < 0.3); lr = rng.normal(size=n)
    x0 = synth.trajectory(W - 1, rng); x0_sb = rng.normal(0, 0.1, 9)
    prior = api.MargPrior.from_arrays(W, LJ, lr, x0, x0_sb)
    assert prior.n == n and prior.W == W
    poses = np.zeros((W, 7)); poses[:, 3] = 1.0
    poses[:W - 1] = synth.perturb(x0, rng, sig_t=0.1, sig_r_deg=3.0)
    poses[W - 1] = [5, 0, 0, 1, 0, 0, 0]
    poses[1, 3:] *= -1.0                                   # q and -q are the same rotation: the w < 0 branch of :246-252
    sb = rng.normal(0, 0.1, (W, 9))
    hf = api.HostFactorSet(); hf.set_marg_prior(prior)
    H, g, c = hf.evaluate(poses, sb)
    prob = oracle.WindowProblem(poses, sb, synth.Q_LB, synth.T_LB)
    prob.set_marg_prior(dict(W=W, lin_jac=LJ, lin_res=lr, x0_pose=x0, x0_sb=x0_sb))
    Ho, go, co = prob.host_normal_eq()
    assert abs(c - co) <= 1e-11 * abs(co)
    assert np.max(np.abs(H - Ho)) <= 1e-10 * np.abs(Ho).max() and np.max(np.abs(g - go)) <= 1e-10 * np.abs(go).max()


@pytest.mark.gpu
def test_integrated_marginalisation_two_windows(oracle):
    """Window 1: associate, solve, marginalise KF0 (device LiDAR blocks with ambient x,y,z columns + host factors -> Schur ->
    eigen) -> prior.  Window 2 (shifted by one keyframe): the prior takes part in the solve and in the next marginalisation.
    Every stage against the oracle's literal restatement: A, b, J^T J, J^T r, keep_block_data, solve iterates."""
    from glio_b200 import api
    W, Q = 5, 1500
    P, sb, spec = _window_with_factors(oracle, W=W + 1, Q=Q, M=40000, seed=101)      # W + 1 keyframes: two windows
    ctx = api.Context(0)
    try:
        ctx.set_map(P["map_xyz"])
        tree = oracle.KdTree(P["map_xyz"])
        prior = None; oprior = None
        for win in range(2):
            ks = list(range(win, win + W))
            poses0 = P["poses_init"][ks]; sb0 = sb[ks]
            s2 = dict(prior=(0,) + spec["prior"][1:] if win == 0 else None, between=[(i - win, j - win) + tuple(rest) for (i, j, *rest) in spec["between"] if win <= i and j < win + W])
            s2["between"] = [tuple(b) for b in s2["between"]]
            hf = api.HostFactorSet()
            if s2["prior"] is not None:
                hf.add_prior(*s2["prior"])
            for bf in s2["between"]:
                hf.add_between(*bf)
            hf.set_marg_prior(prior)
            ctx.window_set_scans([P["scans"][k] for k in ks])
            ctx.window_associate(poses0)
            band = max(29, hf.marg_half_bandwidth())
            rg = ctx.window_solve(poses0, sb0, hf, api.default_solver_options(), band=band)
            # oracle: same matches (association parity is tested elsewhere), same factors
            matches = []
            for i, k in enumerate(ks):
                t2, q2 = ctx.lidar_pose(poses0[i])
                o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][k], t2, q2, tree=tree)
                v = o["status"] == oracle.GO_VALID
                matches.append((P["scans"][k][v], o["nsd"][v], o["score"][v]))
            prob = oracle.WindowProblem(poses0, sb0, P["q_lb"], P["t_lb"], huber_delta=1.0)
            for i, (cp, nsd, score) in enumerate(matches):
                prob.add_unary(np.full(len(cp), i, np.int32), cp, nsd, score)
            if s2["prior"] is not None:
                prob.add_prior(*s2["prior"])
            for bf in s2["between"]:
                prob.add_between(*bf)
            prob.set_marg_prior(oprior)
            ro = prob.solve(oracle.solver_options(), mode=0)
            assert rg["summary"].num_iterations == ro["summary"].num_iterations >= 3
            for a, b in zip(rg["steps"], ro["steps"]):
                a = a.reshape(W, 15); b = b.reshape(W, 15)
                assert np.max(np.abs(a[:, :3] - b[:, :3])) <= 1e-6 and np.max(2 * np.linalg.norm(a[:, 3:6] - b[:, 3:6], axis=1)) <= 1e-8
            # marginalise at the SAME state on both sides (the oracle's optimum), so the comparison isolates the pass itself
            prob.reset_state(ro["poses"], ro["speed_bias"])
            om = prob.marginalize(eps=1e-8, mode=0)
            prior = ctx.window_marginalize(ro["poses"], ro["speed_bias"], hf)
            pa = prior.arrays()
            JtJ_o = om["lin_jac"].T @ om["lin_jac"]; Jtr_o = om["lin_jac"].T @ om["lin_res"]
            assert np.max(np.abs(pa["A_info"] - JtJ_o)) <= 1e-9 * np.abs(JtJ_o).max()
            assert np.max(np.abs(pa["b_info"] - Jtr_o)) <= 1e-8 * np.abs(Jtr_o).max()
            assert np.max(np.abs(pa["lin_jac"].T @ pa["lin_jac"] - pa["A_info"])) <= 1e-12 * np.abs(pa["A_info"]).max()
            assert np.array_equal(pa["x0_pose"], om["x0_pose"]) and np.array_equal(pa["x0_sb"], om["x0_sb"])
            assert prior.W == W and prior.n == 6 * W + 3
            # block-diagonal by construction: unary LiDAR factors + one IMU-like factor -> the next solve keeps its narrow band
            hf2 = api.HostFactorSet(); hf2.set_marg_prior(prior)
            assert hf2.marg_half_bandwidth() <= 14
            oprior = dict(W=W, lin_jac=om["lin_jac"], lin_res=om["lin_res"], x0_pose=om["x0_pose"], x0_sb=om["x0_sb"])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_async_marginalisation_equals_the_synchronous_pass(oracle):
    """glio_window_marginalize_async: the host half runs on the context's worker thread while the caller re-associates (and even
    evaluates) on the same context; the prior is bit-identical to the synchronous pass, twice in a row, and an abandoned job is
    joined by glio_destroy"""
    from glio_b200 import api, synth
    W, Q = 5, 3000
    P = synth.window_problem(W=W, Q=Q, M=60000, seed=41)
    ctx = api.Context(0)
    try:
        ctx.set_map(P["map_xyz"]); ctx.window_set_scans(P["scans"])
        ctx.window_associate(P["poses_init"])
        rng = np.random.default_rng(3); sb = rng.normal(0, 0.05, (W, 9))
        hf = api.HostFactorSet(); T = P["poses_true"]
        sw = np.concatenate([np.full(3, 20.0), np.full(3, 50.0), np.full(9, 5.0)])
        hf.add_prior(0, T[0, :3], T[0, 3:], sb[0], sw)
        for i in range(W - 1):
            dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:]); dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
            hf.add_between(i, i + 1, dp, dq, np.zeros(3), 0.1, sw * 0.5)
        ref = ctx.window_marginalize(P["poses_init"], sb, hf).arrays()
        for _ in range(2):
            job = ctx.window_marginalize_async(P["poses_init"], sb, hf)
            nm = ctx.window_associate(P["poses_init"])               # the GPU is busy with the next association meanwhile
            e = ctx.eval_unary(P["poses_init"])                      # and a new evaluation must not disturb the job's result
            got = job.wait().arrays()
            for k in ("lin_jac", "lin_res", "A_info", "b_info", "x0_pose", "x0_sb"):
                assert np.array_equal(got[k], ref[k]), k
            assert nm.sum() > 0 and np.isfinite(e["cost"]).all()
        with pytest.raises(api.GlioError):
            j1 = ctx.window_marginalize_async(P["poses_init"], sb, hf)
            try:
                ctx.window_marginalize_async(P["poses_init"], sb, hf)     # one job at a time
            finally:
                j1.wait()
        ctx.window_marginalize_async(P["poses_init"], sb, hf)        # abandoned: close() joins it
    finally:
        ctx.close()
