"""CPU tests pinning the oracle.  The reference ships no tests/golden vectors for this path (PARITY UNPINNED), so the
pins are: Ceres 2.0.0's own known-answer tests for the third-party arithmetic (transcribed from the vendored tarball),
analytic planted-plane fixtures, finite differences, and an independent numpy/scipy restatement."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from glio_b200 import synth


# ---------------- Ceres known answers (ceres.tgz::internal/ceres/*_test.cc) ----------------
def test_corrector_scalar_cases(oracle):
    # corrector_test.cc:58-147 ScalarCorrection / ZeroResidual / AlphaClamped
    r = np.sqrt(3.0); J = 10.0
    for rho, alpha in (([3.0, 0.1, -0.01], 0.0), ([3.0, 0.1, -0.1], 0.0)):
        c = oracle.corrector(r * r, rho)
        assert c[2] == 0.0                                    # rho'' < 0 -> clamped branch
        assert c[1] * r == pytest.approx(r * np.sqrt(rho[1]) / (1 - alpha), abs=1e-6)
        assert c[0] * J == pytest.approx(np.sqrt(rho[1]) * (1 - alpha) * J, abs=1e-6)
    c = oracle.corrector(0.0, [0.0, 0.1, -0.01])
    assert c[0] == pytest.approx(np.sqrt(0.1)) and c[1] == pytest.approx(np.sqrt(0.1)) and c[2] == 0.0
    # positive curvature branch (corrector.cc:97-110): D = 1 + 2 s rho''/rho', alpha = 1 - sqrt(D)
    s, rho = 2.0, [2.0, 0.5, 0.05]
    c = oracle.corrector(s, rho)
    D = 1 + 2 * s * rho[2] / rho[1]; alpha = 1 - np.sqrt(D)
    assert c[1] == pytest.approx(np.sqrt(rho[1]) / (1 - alpha)) and c[2] == pytest.approx(alpha / s)


@pytest.mark.parametrize("a", [0.7, 1.3])
@pytest.mark.parametrize("s", [0.357, 1.792])
def test_huber_matches_definition_and_derivatives(oracle, a, s):
    # loss_function_test.cc:86-91 AssertLossFunctionIsValid: rho' and rho'' agree with finite differences of rho
    rho = oracle.huber(a, s)
    h = 1e-6
    f = lambda x: oracle.huber(a, x)[0]
    assert rho[1] == pytest.approx((f(s + h) - f(s - h)) / (2 * h), rel=1e-5)
    assert rho[2] == pytest.approx((oracle.huber(a, s + h)[1] - oracle.huber(a, s - h)[1]) / (2 * h), rel=1e-4, abs=1e-9)
    assert rho[0] == pytest.approx(s if s <= a * a else 2 * a * np.sqrt(s) - a * a)


def _qprod(z, w):
    return np.array([z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3], z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2],
                     z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1], z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0]])


def test_quaternion_parameterization_zero_nearzero_away(oracle):
    # local_parameterization_test.cc:305-352
    x = np.array([0.5, 0.5, 0.5, 0.5])
    assert np.allclose(oracle.quat_plus(x, [0, 0, 0]), x, atol=1e-14)
    x = np.array([0.52, 0.25, 0.15, 0.45]); x /= np.linalg.norm(x)
    d = np.array([0.24, 0.15, 0.10]) * 1e-14
    assert np.allclose(oracle.quat_plus(x, d), _qprod([1.0, *d], x), atol=1e-14)
    d = np.array([0.24, 0.15, 0.10]); nd = np.linalg.norm(d)
    qd = np.array([np.cos(nd), *(np.sin(nd) / nd * d)])
    out = oracle.quat_plus(x, d)
    assert np.allclose(out, _qprod(qd, x), atol=1e-14) and abs(np.linalg.norm(out) - 1) < 1e-14
    # ComputeJacobian == d Plus / d delta at 0 (central differences)
    J = oracle.quat_plus_jacobian(x)
    h = 1e-7
    for k in range(3):
        e = np.zeros(3); e[k] = h
        assert np.allclose((oracle.quat_plus(x, e) - oracle.quat_plus(x, -e)) / (2 * h), J[:, k], atol=1e-8)


# ---------------- planted planes / independent numpy restatement ----------------
def test_plane_solve_planted(oracle):
    rng = np.random.default_rng(0)
    for c in (2.5, -1.8):
        A = np.column_stack([rng.uniform(-3, 3, 5), rng.uniform(-3, 3, 5), np.full(5, c)])
        x, npiv = oracle.plane_solve5(A)
        assert npiv == 3
        n = x / np.linalg.norm(x); d = 1 / np.linalg.norm(x)
        assert np.allclose(n, [0, 0, -np.sign(c)], atol=1e-12) and d == pytest.approx(abs(c), rel=1e-12)
    for _ in range(200):                                                   # vs SVD least squares
        A = rng.normal(size=(5, 3)) * rng.uniform(0.1, 50) + rng.normal(size=3) * 30
        x, _ = oracle.plane_solve5(A)
        ref = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        assert np.allclose(x, ref, rtol=1e-9, atol=1e-12 * np.linalg.norm(ref))


def test_knn_brute_kdtree_scipy_agree(oracle):
    P = synth.window_problem(W=1, Q=3000, M=20000, seed=4)
    m = P["map_xyz"]; q = oracle.transform_points(P["scans"][0], P["poses_init"][0, :3], P["poses_init"][0, 3:])
    ib, db, tie = oracle.knn5_brute(m, q)
    ik, dk = oracle.KdTree(m).knn5(q)
    assert not tie.any(), "synthetic data must be tie-free (jitter)"
    assert np.array_equal(ib, ik) and np.array_equal(db, dk)
    dd, ii = cKDTree(m.astype(np.float64)).query(q.astype(np.float64), k=5)
    assert (np.sort(ii, axis=1) == np.sort(ib, axis=1)).mean() > 0.9999      # sets agree (float32 vs float64 metric)
    assert np.allclose(dd ** 2, db, rtol=1e-5)


def test_transform_matches_rotation_matrix(oracle):
    rng = np.random.default_rng(1)
    p = rng.normal(size=(1000, 3)).astype(np.float32) * 30
    q = synth.quat_from_rpy(0.1, -0.2, 1.3); t = np.array([5.0, -3.0, 0.7])
    out = oracle.transform_points(p, t, q)
    ref = (p.astype(np.float64) @ synth.quat_to_R(q).T + t)
    assert np.max(np.abs(out - ref)) < 1e-5 and out.dtype == np.float32


def test_assoc_numpy_restatement(oracle):
    """Independent restatement of Estimator.cpp:3633-3708 in numpy (float64 lstsq instead of Householder QR)."""
    P = synth.window_problem(W=1, Q=1500, M=20000, seed=8)
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][0, :3], P["poses_init"][0, 3:])
    o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][0], t2, q2)
    m = P["map_xyz"].astype(np.float64)
    n_checked = 0
    for i in range(0, 1500, 7):
        pm = o["pm"][i]
        if o["status"][i] == oracle.GO_FAIL_RADIUS:
            assert not (o["sqd5"][i, 4] < 1.5)
            continue
        A = m[o["idx5"][i]]
        x = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        n = x / np.linalg.norm(x); d = 1 / np.linalg.norm(x)
        assert np.allclose(o["plane"][i], [*n, d], rtol=1e-9)
        valid = (np.abs(A @ n + d) <= 0.18).all()
        if valid:
            pd = np.float32(n @ pm.astype(np.float64) + d)
            w = np.float32(1 - 0.9 * abs(pd) / np.sqrt(np.sqrt(np.float32(pm @ pm))))
            assert abs(w - o["weight"][i]) <= 2e-7
            if w > 0.3:
                assert o["status"][i] == oracle.GO_VALID
                assert np.allclose(o["nsd"][i], np.float32(w) * np.array([*n, d]), rtol=1e-6)
                assert o["score"][i] == pytest.approx(7.5 * float(o["weight"][i]), rel=1e-15)
                n_checked += 1
        else:
            assert o["status"][i] == oracle.GO_FAIL_PLANE
    assert n_checked > 100


def test_weight_gate_branch(oracle):
    """Points far off the plane but within the squared-radius gate must fail the weight gate (Estimator.cpp:3681)."""
    rng = np.random.default_rng(2)
    gx, gy = np.meshgrid(np.arange(-3, 3, 0.1), np.arange(-3, 3, 0.1))
    m = np.column_stack([gx.ravel() + 10, gy.ravel() + 10, np.full(gx.size, 0.5)]).astype(np.float32)
    m += rng.uniform(-1e-3, 1e-3, m.shape).astype(np.float32)
    scan = np.array([[10.0, 10.0, 0.5 + h] for h in (0.0, 0.3, 0.9, 1.1)], np.float32)
    o = oracle.assoc_scan_to_map(m, scan, [0, 0, 0], [1, 0, 0, 0])
    # range term: sqrt(sqrt(|p|^2)) = sqrt(|p|) ~ 3.76 ; weight = 1 - 0.9 h / 3.76
    assert o["status"][0] == oracle.GO_VALID and o["status"][1] == oracle.GO_VALID
    assert o["status"][2] == oracle.GO_VALID             # sqd5 ~ 0.81+ < 1.5 and weight ~ 0.78
    assert o["status"][3] == oracle.GO_VALID or o["status"][3] == oracle.GO_FAIL_RADIUS
    far = np.array([[0.4, 0.3, 0.5 + 1.0]], np.float32)       # close to the world origin -> tiny range term -> weight < 0.3
    m2 = (m - np.array([10, 10, 0], np.float32))
    o2 = oracle.assoc_scan_to_map(m2, far, [0, 0, 0], [1, 0, 0, 0])
    assert o2["status"][0] == oracle.GO_FAIL_WEIGHT and o2["weight"][0] < 0.3


# ---------------- factors: closed form == Jet autodiff == finite differences ----------------
def _fd_tangent(fun, pose, oracle, h=1e-6):
    J = np.zeros(6)
    for k in range(6):
        d = np.zeros(6); d[k] = h
        pp = pose.copy(); pm = pose.copy()
        pp[:3] += d[:3]; pm[:3] -= d[:3]
        pp[3:] = oracle.quat_plus(pose[3:], d[3:]); pm[3:] = oracle.quat_plus(pose[3:], -d[3:])
        J[k] = (fun(pp) - fun(pm)) / (2 * h)
    return J


def test_unary_factor_jacobians(oracle):
    rng = np.random.default_rng(3)
    pose = np.array([1.0, -2.0, 0.3, *synth.quat_from_rpy(0.2, -0.1, 0.7)])
    q_lb = synth.quat_from_rpy(0.01, 0.02, -0.03); t_lb = np.array([0.1, -0.05, 0.28])
    N = 50
    cp = rng.normal(size=(N, 3)).astype(np.float32) * 10; nsd = rng.normal(size=(N, 4)).astype(np.float32); score = rng.uniform(1, 7, N)
    kf = np.zeros(N, np.int32)
    a = oracle.eval_unary(pose[None], q_lb, t_lb, kf, cp, nsd, score, huber_delta=0.0, mode=0)
    b = oracle.eval_unary(pose[None], q_lb, t_lb, kf, cp, nsd, score, huber_delta=0.0, mode=1)
    assert np.allclose(a["r"], b["r"], rtol=1e-13) and np.allclose(a["J"], b["J"], rtol=1e-11, atol=1e-11)
    for i in range(0, N, 7):
        f = lambda p: oracle.eval_unary(p[None], q_lb, t_lb, kf[i:i + 1], cp[i:i + 1], nsd[i:i + 1], score[i:i + 1], huber_delta=0.0)["r"][0]
        assert np.allclose(_fd_tangent(f, pose, oracle), a["J"][i], rtol=1e-6, atol=1e-6)
    # Huber: r and J are scaled by sqrt(rho') in the outlier region, cost = 0.5 rho
    h = oracle.eval_unary(pose[None], q_lb, t_lb, kf, cp, nsd, score, huber_delta=1.0, mode=0)
    out = np.abs(a["r"]) > 1
    assert out.any()
    sc = np.where(out, 1 / np.sqrt(np.abs(a["r"])), 1.0)
    assert np.allclose(h["r"], a["r"] * sc, rtol=1e-13) and np.allclose(h["J"], a["J"] * sc[:, None], rtol=1e-13)
    assert np.allclose(h["cost"], np.where(out, 0.5 * (2 * np.abs(a["r"]) - 1), 0.5 * a["r"] ** 2), rtol=1e-13)
    # marginalisation variant: ambient x,y,z columns (MarginalizationFactor.cpp:9-12) differ from the tangent ones
    mj = oracle.eval_unary(pose[None], q_lb, t_lb, kf, cp, nsd, score, huber_delta=0.0, mode=0, jac_kind=1)
    mc = oracle.eval_unary(pose[None], q_lb, t_lb, kf, cp, nsd, score, huber_delta=0.0, mode=1, jac_kind=1)
    assert np.allclose(mj["J"], mc["J"], rtol=1e-11, atol=1e-11) and not np.allclose(mj["J"][:, 3:], a["J"][:, 3:], rtol=1e-3)


def test_binary_and_edge_factor_jacobians(oracle):
    rng = np.random.default_rng(4)
    poses = np.array([[1.0, -2.0, 0.3, *synth.quat_from_rpy(0.2, -0.1, 0.7)], [2.0, -1.5, 0.2, *synth.quat_from_rpy(-0.1, 0.05, 0.9)]])
    N = 40
    cp = rng.normal(size=(N, 3)).astype(np.float32) * 8
    nc = rng.normal(size=(N, 6)); nc[:, :3] /= np.linalg.norm(nc[:, :3], axis=1)[:, None]
    score = rng.uniform(0.5, 2.5, N); kc = np.zeros(N, np.int32); ko = np.ones(N, np.int32)
    a = oracle.eval_binary(poses, kc, ko, cp, nc, score, mode=0); b = oracle.eval_binary(poses, kc, ko, cp, nc, score, mode=1)
    assert np.allclose(a["r"], b["r"], rtol=1e-13) and np.allclose(a["J"], b["J"], rtol=1e-10, atol=1e-10)
    assert np.allclose(a["H"], b["H"], rtol=1e-10)
    pa = rng.normal(size=(N, 3)).astype(np.float32) * 5; pb = pa + rng.normal(size=(N, 3)).astype(np.float32)
    s = rng.uniform(0.5, 3, N); kf = np.zeros(N, np.int32)
    q_lb = [1, 0, 0, 0]; t_lb = [0, 0, 0.28]
    e0 = oracle.eval_edge(poses[:1], q_lb, t_lb, kf, cp, pa, pb, s, huber_delta=0.0, mode=0)
    e1 = oracle.eval_edge(poses[:1], q_lb, t_lb, kf, cp, pa, pb, s, huber_delta=0.0, mode=1)
    assert np.allclose(e0["r"], e1["r"], rtol=1e-13) and np.allclose(e0["J"], e1["J"], rtol=1e-9, atol=1e-9)


# ---------------- host factors (product, analytic) vs oracle (Jets) — CPU only ----------------
def test_host_factors_match_oracle_autodiff(oracle):
    from glio_b200 import api
    rng = np.random.default_rng(6)
    W = 4
    P = synth.window_problem(W=W, Q=10, M=100, seed=3)
    sb = rng.normal(0, 0.2, (W, 9))
    for use_sb in (False, True):
        prob = oracle.WindowProblem(P["poses_init"], sb if use_sb else None, P["q_lb"], P["t_lb"])
        hf = api.HostFactorSet()
        sw = np.abs(rng.normal(3, 1, 15))
        a = (1, P["poses_true"][1, :3], P["poses_true"][1, 3:], sb[1], sw); prob.add_prior(*a); hf.add_prior(*a)
        for i in range(W - 1):
            a = (i, i + 1, rng.normal(size=3), synth.quat_from_rpy(*rng.normal(0, 0.1, 3)), rng.normal(size=3), 0.1, sw * 2)
            prob.add_between(*a); hf.add_between(*a)
        a = (2, [0.1, 0.2, 0.3], [100.0, 200.0, 300.0], 370.0, 0.5); prob.add_range(*a); hf.add_range(*a)
        Ho, go_, co = prob.host_normal_eq()
        Hp, gp, cp_ = hf.evaluate(P["poses_init"], sb if use_sb else None)
        assert np.max(np.abs(Ho - Hp)) <= 1e-12 * np.max(np.abs(Ho)) and np.max(np.abs(go_ - gp)) <= 1e-12 * np.max(np.abs(go_))
        assert cp_ == pytest.approx(co, rel=1e-13)


def test_oracle_window_solve_converges(oracle):
    P = synth.window_problem(W=3, Q=800, M=20000, seed=12)
    tree = oracle.KdTree(P["map_xyz"])
    prob = oracle.WindowProblem(P["poses_init"], None, P["q_lb"], P["t_lb"])
    for k in range(3):
        t2, q2 = synth.lidar_pose_in_world(P["poses_init"][k, :3], P["poses_init"][k, 3:])
        o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][k], t2, q2, tree=tree)
        v = o["status"] == 0
        prob.add_unary(np.full(v.sum(), k, np.int32), P["scans"][k][v], o["nsd"][v], o["score"][v])
    r0 = prob.solve(oracle.solver_options(), mode=0)
    prob.reset_state(P["poses_init"])
    r1 = prob.solve(oracle.solver_options(), mode=1)            # closed-form Jacobians give the same iterations
    assert r0["summary"].termination == 0 and r0["summary"].final_cost < 0.5 * r0["summary"].initial_cost
    assert r0["summary"].num_iterations == r1["summary"].num_iterations
    assert np.allclose(r0["poses"], r1["poses"], atol=1e-9)
    costs = [it["cost"] for it in r0["iterations"] if it["step_is_successful"]]
    assert all(b < a for a, b in zip(costs, costs[1:]))       # monotone (use_nonmonotonic_steps = false)
