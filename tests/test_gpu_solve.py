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
