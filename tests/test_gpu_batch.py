"""K1b + K2b parity: scan-to-multiscan pair association and binary plane factor blocks vs the CPU oracle."""
import numpy as np
import pytest

from glio_b200 import synth

pytestmark = pytest.mark.gpu
REL = 1e-11


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def setup(oracle):
    from glio_b200 import api
    K, Q, sr = 6, 4000, 2
    B = synth.batch_problem(K=K, Q=Q, seed=31, search_range=sr, rng_range=5.0)
    ctx = api.Context(0)
    for k in range(K):
        ctx.batch_set_frame(k, B["scans"][k], B["poses_init"][k])
    cur, oth = [], []
    for i in range(K):
        for j in range(max(0, i - sr), min(K, i + sr + 1)):
            if j != i:
                cur.append(i); oth.append(j)
    nm = ctx.batch_associate_pairs(cur, oth)
    yield ctx, B, np.array(cur), np.array(oth), nm
    ctx.close()


def test_pair_association_bit_exact(setup, oracle):
    ctx, B, cur, oth, nm = setup
    Q = B["Q"]
    for i, (c, o) in enumerate(zip(cur, oth)):
        if i % 3:      # a third of the pairs is plenty
            continue
        ro = oracle.assoc_pair(B["scans"][c], B["poses_init"][c, :3], B["poses_init"][c, 3:], B["scans"][o],
                               B["poses_init"][o, :3], B["poses_init"][o, 3:])
        m = ctx.batch_get_matches(int(c), int(o), Q)
        v = ro["status"] == oracle.GO_VALID
        assert m["n"] == nm[i] == ro["nvalid"] == int(v.sum()) and m["n"] > 100
        assert np.array_equal(m["src"], np.nonzero(v)[0].astype(np.int32))
        assert np.array_equal(m["cp"], B["scans"][c][v])
        assert np.array_equal(m["weight"], ro["weight"][v])
        assert np.array_equal(m["normal_cent"], ro["normal_cent"][v]), "local-frame normal/centroid must be bit-exact"


def _oracle_blocks(ctx, oracle, B, cur, oth, poses, sel=None):
    kc, ko, cp, nc, sc = [], [], [], [], []
    for i, (c, o) in enumerate(zip(cur, oth)):
        m = ctx.batch_get_matches(int(c), int(o), B["Q"])
        idx = np.arange(m["n"]) if sel is None else sel[i]
        kc.append(np.full(len(idx), c, np.int32)); ko.append(np.full(len(idx), o, np.int32))
        cp.append(m["cp"][idx]); nc.append(m["normal_cent"][idx]); sc.append(ctx.params.batch_score * m["weight"][idx].astype(np.float64))
    return oracle.eval_binary(poses, np.concatenate(kc), np.concatenate(ko), np.concatenate(cp), np.concatenate(nc),
                              np.concatenate(sc), huber_delta=0.0, mode=0, per_residual=False)


@pytest.mark.parametrize("which", ["init", "true"])
def test_binary_blocks(setup, oracle, which):
    ctx, B, cur, oth, nm = setup
    poses = B["poses_init"] if which == "init" else B["poses_true"]
    o = _oracle_blocks(ctx, oracle, B, cur, oth, poses)
    g = ctx.eval_binary(poses)
    pc, po_ = ctx.batch_pair_list()
    assert sorted(zip(pc.tolist(), po_.tolist())) == sorted(zip(cur.tolist(), oth.tolist()))   # internal order: grouped by searched frame
    K = B["K"]
    for k in range(K):
        assert _rel(g["Hdiag"][k], o["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6]) < REL
        assert _rel(g["g"][k], o["g"][6 * k:6 * k + 6]) < REL
    # off-diagonal: the oracle's dense H block (cur,oth) sums the pair (cur,oth) AND the transpose of pair (oth,cur)
    dense = np.zeros((6 * K, 6 * K))
    for p, (c, oo) in enumerate(zip(pc, po_)):
        dense[6 * c:6 * c + 6, 6 * oo:6 * oo + 6] += g["Hoff"][p]
        dense[6 * oo:6 * oo + 6, 6 * c:6 * c + 6] += g["Hoff"][p].T
    for k in range(K):
        dense[6 * k:6 * k + 6, 6 * k:6 * k + 6] = g["Hdiag"][k]
    assert _rel(dense, o["H"]) < REL
    assert abs(g["cost"] - o["cost_total"]) <= REL * o["cost_total"]
    c2 = ctx.eval_binary(poses, want_jac=False)["cost"]
    assert abs(c2 - g["cost"]) <= 1e-13 * g["cost"]


def test_batch_selection(setup, oracle):
    ctx, B, cur, oth, nm = setup
    sel = [np.arange(min(25, n), dtype=np.int32) for n in nm]          # faithful mode: 25 per pair (batch_feature_res_num)
    for (c, o), s in zip(zip(cur, oth), sel):
        ctx.batch_select(int(c), int(o), s)
    try:
        o = _oracle_blocks(ctx, oracle, B, cur, oth, B["poses_init"], sel)
        g = ctx.eval_binary(B["poses_init"])
        for k in range(B["K"]):
            assert _rel(g["Hdiag"][k], o["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6]) < REL
        assert abs(g["cost"] - o["cost_total"]) <= REL * o["cost_total"]
    finally:
        for c, o_ in zip(cur, oth):
            ctx.batch_select(int(c), int(o_), None)


def test_batch_solve_matches_oracle(setup, oracle):
    """optimizeBatch-style solve (SUBSPACE_DOGLEG, nonmonotonic) over binary plane factors + an IMU-like chain/prior."""
    from glio_b200 import api
    ctx, B, cur, oth, nm = setup
    K = B["K"]; T = B["poses_true"]
    rng = np.random.default_rng(5)
    prob = oracle.WindowProblem(B["poses_init"], None, [1, 0, 0, 0], [0, 0, 0], huber_delta=0.0)
    for c, o in zip(cur, oth):
        m = ctx.batch_get_matches(int(c), int(o), B["Q"])
        prob.add_binary(np.full(m["n"], c, np.int32), np.full(m["n"], o, np.int32), m["cp"], m["normal_cent"],
                        ctx.params.batch_score * m["weight"].astype(np.float64))
    hf = api.HostFactorSet()
    sw = np.concatenate([np.full(3, 10.0), np.full(3, 30.0), np.zeros(9)])
    a = (0, T[0, :3], T[0, 3:], None, sw * 3)
    prob.add_prior(*a); hf.add_prior(*a)
    for i in range(K - 1):
        dq = synth.quat_mul(synth.quat_conj(T[i, 3:]), T[i + 1, 3:])
        dp = synth.quat_to_R(T[i, 3:]).T @ (T[i + 1, :3] - T[i, :3])
        a = (i, i + 1, dp + rng.normal(0, 0.005, 3), dq, np.zeros(3), 0.1, sw)
        prob.add_between(*a); hf.add_between(*a)
    _compare_batch_solves(ctx, oracle, api, B, prob, hf, K, 1e4)
    # a small initial radius forces the subspace-dogleg boundary branch (quartic roots) and some rejected steps
    _compare_batch_solves(ctx, oracle, api, B, prob, hf, K, 0.02)


def _compare_batch_solves(ctx, oracle, api, B, prob_template, hf, K, radius0):
    import copy
    oo = oracle.solver_options(dogleg_type=1, use_nonmonotonic_steps=1, max_num_iterations=30, initial_trust_region_radius=radius0)
    prob = prob_template
    prob.reset_state(B["poses_init"])
    ro = prob.solve(oo, mode=0)
    rg = ctx.batch_solve(B["poses_init"], None, hf, api.batch_solver_options(max_num_iterations=30, initial_trust_region_radius=radius0))
    so, sg = ro["summary"], rg["summary"]
    assert sg.termination == so.termination and sg.message == so.message and sg.num_iterations == so.num_iterations >= 3
    assert len(rg["steps"]) == len(ro["steps"])
    for a_, b_ in zip(rg["steps"], ro["steps"]):
        a_ = a_.reshape(K, 6); b_ = b_.reshape(K, 6)
        assert np.max(np.abs(a_[:, :3] - b_[:, :3])) <= 1e-6
        assert np.max(2 * np.linalg.norm(a_[:, 3:] - b_[:, 3:], axis=1)) <= 1e-8
    for ig, io in zip(rg["iterations"], ro["iterations"]):
        assert ig["step_is_successful"] == io["step_is_successful"]
        assert ig["cost"] == pytest.approx(io["cost"], rel=1e-9)
    assert np.max(np.abs(rg["poses"] - ro["poses"])) <= 1e-6
    assert sg.final_cost < sg.initial_cost


def test_uploaded_pair_matches_equal_associated_ones(setup):
    """glio_batch_set_pair_matches (the path the Ceres shim takes for BinaryLidarPlaneNormFactor blocks): a second
    context fed with the downloaded matches of the first evaluates to the same blocks."""
    from glio_b200 import api
    ctx, B, cur, oth, nm = setup
    pc, po_ = ctx.batch_pair_list()
    g1 = ctx.eval_binary(B["poses_init"])
    c2 = api.Context(0)
    try:
        for c, o in zip(pc, po_):
            m = ctx.batch_get_matches(int(c), int(o), B["Q"])
            c2.batch_set_pair_matches(int(c), int(o), m["cp"], m["normal_cent"], m["weight"])
        p2c, p2o = c2.batch_pair_list()
        assert np.array_equal(p2c, pc) and np.array_equal(p2o, po_)
        g2 = c2.eval_binary(B["poses_init"])
        for k in ("Hdiag", "Hoff", "g"):
            assert np.allclose(g2[k], g1[k], rtol=1e-13, atol=0), k
        assert abs(g2["cost"] - g1["cost"]) <= 1e-13 * g1["cost"]
    finally:
        c2.close()


def test_pair_association_at_cfg3_scan_size(oracle):
    """One scan-to-multiscan pair at the BASELINE cfg 3 / cfg 4 scan size (Q = 100 k points per keyframe, Estimator.cpp:3710-3806):
    the compacted match list, weights and local-frame normal / centroid bit-exact against the oracle (kd-tree over the searched
    frame, all host threads)."""
    from glio_b200 import api
    B = synth.batch_problem(K=13, Q=100_000, seed=synth.SEED0 + 3, search_range=6, frames=[5, 8])
    ctx = api.Context(0)
    try:
        for k in (5, 8):
            ctx.batch_set_frame(k, B["scans"][k], B["poses_init"][k])
        nm = ctx.batch_associate_pairs([5, 8], [8, 5])
        for i, (c, o) in enumerate(((5, 8), (8, 5))):
            ro = oracle.assoc_pair(B["scans"][c], B["poses_init"][c, :3], B["poses_init"][c, 3:], B["scans"][o],
                                   B["poses_init"][o, :3], B["poses_init"][o, 3:])
            m = ctx.batch_get_matches(c, o, 100_000)
            v = ro["status"] == oracle.GO_VALID
            assert m["n"] == nm[i] == ro["nvalid"] == int(v.sum()) and m["n"] > 50_000
            assert np.array_equal(m["src"], np.nonzero(v)[0].astype(np.int32))
            assert np.array_equal(m["weight"], ro["weight"][v])
            assert np.array_equal(m["normal_cent"], ro["normal_cent"][v])
    finally:
        ctx.close()
