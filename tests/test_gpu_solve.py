"""Window solve parity: the device-backed minimizer vs the CPU oracle's literal Ceres restatement.
Per-iteration tangent updates must agree to <= 1e-6 m / 1e-8 rad (north_star), with identical accept/reject
decisions and trust-region radii."""
import numpy as np
import pytest

from glio_b200 import synth

pytestmark = pytest.mark.gpu


def _build(oracle, api, W, Q, M, seed, use_sb, n_sel=None):
    P = synth.window_problem(W=W, Q=Q, M=M, seed=seed)
    ctx = api.Context(0)
    ctx.set_map(P["map_xyz"])
    ctx.window_set_scans(P["scans"])
    ctx.window_associate(P["poses_init"])
    rng = np.random.default_rng(seed)
    sb0 = rng.normal(0, 0.1, (W, 9)) if use_sb else None
    prob = oracle.WindowProblem(P["poses_init"], sb0, P["q_lb"], P["t_lb"], huber_delta=1.0)
    for k in range(W):
        m = ctx.get_matches(k, Q)
        idx = np.arange(m["n"])
        if n_sel is not None:                      # faithful mode: a fixed index list (first n_sel valid matches)
            idx = idx[:n_sel].astype(np.int32)
            ctx.select(k, idx)
        prob.add_unary(np.full(len(idx), k, np.int32), m["cp"][idx], m["nsd"][idx],
                       ctx.params.lidar_const * m["weight"][idx].astype(np.float64))
    hf = api.HostFactorSet()
    T = P["poses_true"]
    sw = np.concatenate([np.full(3, 20.0), np.full(3, 50.0), np.full(9, 5.0)])
    a = (0, T[0, :3] + 0.01, T[0, 3:], None if sb0 is None else sb0[0], sw)
    prob.add_prior(*a); hf.add_prior(*a)
    for i in range(W - 1):
        dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:])
        dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
        a = (i, i + 1, dp + rng.normal(0, 0.01, 3), dq, np.zeros(3), 0.1, sw * 0.5)
        prob.add_between(*a); hf.add_between(*a)
    for k in range(W):
        sat = np.array([2.0e4 * np.cos(k), 2.0e4 * np.sin(k), 2.0e4])
        rho = np.linalg.norm(T[k, :3] - sat) + 0.3
        a = (k, [0.0, 0.0, 0.0], sat, rho, 0.7)
        prob.add_range(*a); hf.add_range(*a)
    return P, ctx, prob, hf, sb0


@pytest.mark.parametrize("use_sb,n_sel,fuse", [(True, 100, 1), (False, None, 1), (True, None, 0)])
def test_window_solve_matches_oracle(oracle, use_sb, n_sel, fuse):
    from glio_b200 import api
    W = 5
    P, ctx, prob, hf, sb0 = _build(oracle, api, W, 1000, 50000, synth.SEED0 + 1, use_sb, n_sel)
    try:
        ro = prob.solve(oracle.solver_options(), mode=0)
        rg = ctx.window_solve(P["poses_init"], sb0, hf, api.default_solver_options(fuse_candidate_jacobian=fuse))
        so, sg = ro["summary"], rg["summary"]
        assert sg.termination == so.termination and sg.message == so.message
        assert sg.num_iterations == so.num_iterations and sg.num_iterations >= 3
        nt = 15 if use_sb else 6
        assert len(rg["steps"]) == len(ro["steps"]) >= 2
        for a, b in zip(rg["steps"], ro["steps"]):
            a = a.reshape(W, nt); b = b.reshape(W, nt)
            assert np.max(np.abs(a[:, :3] - b[:, :3])) <= 1e-6            # metres
            assert np.max(2 * np.linalg.norm(a[:, 3:6] - b[:, 3:6], axis=1)) <= 1e-8   # rad (delta is a half-angle)
        for ig, io in zip(rg["iterations"], ro["iterations"]):
            assert ig["step_is_successful"] == io["step_is_successful"] and ig["step_is_valid"] == io["step_is_valid"]
            assert ig["trust_region_radius"] == pytest.approx(io["trust_region_radius"], rel=1e-9)
            assert ig["cost"] == pytest.approx(io["cost"], rel=1e-10)
        assert np.max(np.abs(rg["poses"] - ro["poses"])) <= 1e-6
        # it actually optimised something
        assert sg.final_cost < 0.5 * sg.initial_cost
        # band-storage variant of the host callback: same iterates
        rb = ctx.window_solve(P["poses_init"], sb0, hf, api.default_solver_options(fuse_candidate_jacobian=fuse), band=29 if use_sb else 11)
        assert rb["summary"].num_iterations == sg.num_iterations
        assert np.max(np.abs(rb["poses"] - rg["poses"])) <= 1e-12
    finally:
        ctx.close()


def test_front_end_scan_matcher_settings(oracle):
    """the front end's scan-to-map matcher (LidarOdometry.cpp:343-404, :474-540) is the same kernel recipe with
    other constants: squared radius 1.0, plane threshold 0.06, weight gate 0.4, Huber 0.1, no extrinsic, residual
    (w n).(q p + t) + w d without the lidar_const*weight score (LidarPlaneNormIncreFactor) -> unit_score = 1, lidar_const = 1."""
    from glio_b200 import api
    P = synth.window_problem(W=1, Q=4000, M=60000, seed=synth.SEED0 + 9)
    ident_q, zero_t = [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    ctx = api.Context(0, keep_debug=1, kd_max_radius=1.0, surf_dist_thres=0.06, weight_min=0.4, lidar_const=1.0, huber_delta=0.1,
                      q_lb=ident_q, t_lb=zero_t, unit_score=1)
    try:
        ctx.set_map(P["map_xyz"])
        pose = P["poses_init"][0].copy()
        # the scan was generated in the lidar frame of the synthetic extrinsic; with an identity extrinsic the lidar pose IS the state
        t2, q2 = synth.lidar_pose_in_world(pose[:3], pose[3:7])
        state = np.concatenate([t2, q2])[None, :]
        ctx.window_set_scans([P["scans"][0]])
        nm = ctx.window_associate(state)
        prm = oracle.default_params(); prm.kd_max_radius = 1.0; prm.surf_dist_thres = 0.06; prm.weight_min = 0.4; prm.lidar_const = 1.0
        o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][0], t2, q2, prm=prm)
        d = ctx.get_assoc_debug(0, 4000)
        assert np.array_equal(d["status"], o["status"]) and nm[0] == o["nvalid"] > 300
        assert (o["status"] == oracle.GO_FAIL_WEIGHT).any() or (o["status"] == oracle.GO_FAIL_PLANE).any()      # the tighter gates bite
        v = o["status"] == oracle.GO_VALID
        kf = np.zeros(int(v.sum()), np.int32); ones = np.ones(int(v.sum()))
        oe = oracle.eval_unary(state, ident_q, zero_t, kf, P["scans"][0][v], o["nsd"][v], ones, huber_delta=0.1, mode=0)
        ge = ctx.eval_unary(state)
        assert np.max(np.abs(ge["H"][0] - oe["H"])) <= 1e-11 * np.max(np.abs(oe["H"]))
        assert np.max(np.abs(ge["g"][0] - oe["g"])) <= 1e-11 * np.max(np.abs(oe["g"])) and abs(ge["cost"][0] - oe["cost_total"]) <= 1e-11 * oe["cost_total"]
        # one Ceres solve of the front end's loop (6-dof, LiDAR residuals only) with the front end's own options
        # (LidarOdometry.cpp:521-530): trust_region_strategy_type = LEVENBERG_MARQUARDT (Ceres' default), DENSE_QR, at most
        # 4 iterations.  Oracle: LM over an unpivoted Householder QR of [J; D] (reserved = 2); product: LM over the normal equations.
        prob = oracle.WindowProblem(state, None, ident_q, zero_t, huber_delta=0.1)
        prob.add_unary(kf, P["scans"][0][v], o["nsd"][v], ones)
        for max_it in (4, 15):
            prob.reset_state(state)
            ro = prob.solve(oracle.solver_options(reserved=2, max_num_iterations=max_it), mode=0)
            rg = ctx.window_solve(state, None, None, api.default_solver_options(trust_region_strategy=1, max_num_iterations=max_it))
            assert rg["summary"].num_iterations == ro["summary"].num_iterations >= 2
            assert len(rg["steps"]) == len(ro["steps"])
            for a, b in zip(rg["steps"], ro["steps"]):
                assert np.max(np.abs(a[:3] - b[:3])) <= 1e-6 and np.max(np.abs(a[3:6] - b[3:6])) <= 1e-8
            for ig, io in zip(rg["iterations"], ro["iterations"]):
                assert ig["step_is_successful"] == io["step_is_successful"] and ig["step_is_valid"] == io["step_is_valid"]
                assert ig["trust_region_radius"] == pytest.approx(io["trust_region_radius"], rel=1e-7)
            assert np.max(np.abs(rg["poses"] - ro["poses"])) <= 1e-6
            assert rg["summary"].final_cost < 0.7 * rg["summary"].initial_cost
        # and the dogleg solve still agrees as before (the Estimator's strategy on the same problem)
        prob.reset_state(state)
        ro = prob.solve(oracle.solver_options(), mode=0)
        rg = ctx.window_solve(state, None, None, api.default_solver_options())
        assert rg["summary"].num_iterations == ro["summary"].num_iterations >= 2
        for a, b in zip(rg["steps"], ro["steps"]):
            assert np.max(np.abs(a[:3] - b[:3])) <= 1e-6 and np.max(np.abs(a[3:6] - b[3:6])) <= 1e-8
    finally:
        ctx.close()
