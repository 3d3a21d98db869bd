"""K2 parity: CUDA unary plane residual / Jacobian / block normal equations vs the CPU oracle
(Jet autodiff of the reference functor -> QuaternionParameterization -> Huber corrector)."""
import numpy as np
import pytest

from glio_b200 import synth

pytestmark = pytest.mark.gpu

REL = 1e-11   # SURVEY 8(c) "H, g, cost <= 1e-11 relative"


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def setup(oracle):
    from glio_b200 import api
    ctx = api.Context(0)
    P = synth.window_problem(W=5, Q=3000, M=50000, seed=21)
    ctx.set_map(P["map_xyz"])
    ctx.window_set_scans(P["scans"])
    nm = ctx.window_associate(P["poses_init"])
    matches = [ctx.get_matches(k, 3000) for k in range(5)]
    yield ctx, P, nm, matches
    ctx.close()


def _oracle_inputs(ctx, matches, sel=None):
    kf, cp, nsd, score = [], [], [], []
    for k, m in enumerate(matches):
        idx = np.arange(m["n"]) if sel is None else sel[k]
        kf.append(np.full(len(idx), k, np.int32)); cp.append(m["cp"][idx]); nsd.append(m["nsd"][idx])
        score.append(ctx.params.lidar_const * m["weight"][idx].astype(np.float64))
    return np.concatenate(kf), np.concatenate(cp), np.concatenate(nsd), np.concatenate(score)


@pytest.mark.parametrize("jac_kind", [0, 1])
@pytest.mark.parametrize("which", ["init", "true"])
def test_unary_blocks(setup, oracle, jac_kind, which):
    ctx, P, nm, matches = setup
    poses = P["poses_init"] if which == "init" else P["poses_true"]
    kf, cp, nsd, score = _oracle_inputs(ctx, matches)
    o = oracle.eval_unary(poses, P["q_lb"], P["t_lb"], kf, cp, nsd, score, huber_delta=1.0, mode=0, jac_kind=jac_kind)
    g = ctx.eval_unary(poses, jac_kind=jac_kind)
    W = len(poses)
    for k in range(W):
        Hk = o["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6]
        assert _rel(g["H"][k], Hk) < REL
        assert _rel(g["g"][k], o["g"][6 * k:6 * k + 6]) < REL
        ck = o["cost"][kf == k].sum()
        assert abs(g["cost"][k] - ck) <= REL * abs(ck)
    # cost-only evaluation returns the same cost
    c2 = ctx.eval_unary(poses, want_jac=False)["cost"]
    assert np.allclose(c2, g["cost"], rtol=1e-13, atol=0)
    # some residuals must be in the Huber outlier region at the perturbed poses, or the test is vacuous
    if which == "init":
        assert (np.abs(o["r"]) > 0).all() and (o["cost"] > 0.5).any()


def test_unary_per_residual(setup, oracle):
    ctx, P, nm, matches = setup
    for jac_kind in (0, 1):
        for k in (0, 3):
            m = matches[k]
            r, J = ctx.eval_unary_residuals(k, P["poses_init"][k], m["n"], jac_kind=jac_kind)
            score = ctx.params.lidar_const * m["weight"].astype(np.float64)
            o = oracle.eval_unary(P["poses_init"][k:k + 1], P["q_lb"], P["t_lb"], np.zeros(m["n"], np.int32), m["cp"], m["nsd"],
                                  score, huber_delta=1.0, mode=0, jac_kind=jac_kind)
            assert np.max(np.abs(r - o["r"])) <= 1e-12 * max(1.0, np.max(np.abs(o["r"])))
            assert np.max(np.abs(J - o["J"])) <= 1e-12 * max(1.0, np.max(np.abs(o["J"])))


def test_selection_index_list(setup, oracle):
    ctx, P, nm, matches = setup
    rng = np.random.default_rng(3)
    sel = [np.sort(rng.choice(m["n"], size=100, replace=False)).astype(np.int32) for m in matches]
    for k, s in enumerate(sel):
        ctx.select(k, s)
    try:
        kf, cp, nsd, score = _oracle_inputs(ctx, matches, sel)
        o = oracle.eval_unary(P["poses_init"], P["q_lb"], P["t_lb"], kf, cp, nsd, score, huber_delta=1.0, mode=0)
        g = ctx.eval_unary(P["poses_init"])
        for k in range(5):
            assert _rel(g["H"][k], o["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6]) < REL
            assert _rel(g["g"][k], o["g"][6 * k:6 * k + 6]) < REL
        with pytest.raises(Exception):
            ctx.select(0, np.array([10 ** 6], np.int32))
    finally:
        for k in range(5):
            ctx.select(k, None)
    g2 = ctx.eval_unary(P["poses_init"])
    assert g2["cost"].sum() > g["cost"].sum()


def test_edge_blocks(oracle):
    """K2e: LidarEdgeFactor blocks vs the oracle's Jet autodiff of the functor (correspondences are an input)."""
    from glio_b200 import api
    rng = np.random.default_rng(9)
    W, n = 3, 700
    ctx = api.Context(0)
    try:
        P = synth.window_problem(W=W, Q=100, M=1000, seed=2)
        poses = P["poses_init"]
        kf, cp, pa, pb, s = [], [], [], [], []
        for k in range(W):
            a = rng.uniform(-20, 20, (n, 3)); d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
            b = a + d * rng.uniform(0.2, 1.0, (n, 1))
            # scan point: near the line in the world, mapped into the lidar frame of keyframe k (noise -> some Huber outliers)
            pw = a + d * rng.uniform(-1, 2, (n, 1)) + rng.normal(0, 0.08, (n, 3))
            pbody = (pw - poses[k, :3]) @ synth.quat_to_R(poses[k, 3:])
            pl = pbody @ synth.quat_to_R(P["q_lb"]).T + P["t_lb"]
            ss = rng.uniform(1.0, 8.0, n)
            ctx.set_edges(k, pl, a, b, ss)
            kf.append(np.full(n, k, np.int32)); cp.append(pl.astype(np.float32)); pa.append(a.astype(np.float32)); pb.append(b.astype(np.float32))
            s.append(ss.astype(np.float32).astype(np.float64))       # the weight rides in a float32 lane on the device
        kf, cp, pa, pb, s = map(np.concatenate, (kf, cp, pa, pb, s))
        o = oracle.eval_edge(poses, P["q_lb"], P["t_lb"], kf, cp, pa, pb, s, huber_delta=1.0, mode=0)
        oc = oracle.eval_edge(poses, P["q_lb"], P["t_lb"], kf, cp, pa, pb, s, huber_delta=1.0, mode=1)
        assert _rel(oc["H"], o["H"]) < 1e-10        # closed form == autodiff inside the oracle
        g = ctx.eval_edge(poses)
        for k in range(W):
            assert _rel(g["H"][k], o["H"][6 * k:6 * k + 6, 6 * k:6 * k + 6]) < REL
            assert _rel(g["g"][k], o["g"][6 * k:6 * k + 6]) < REL
            ck = o["cost"][kf == k].sum()
            assert abs(g["cost"][k] - ck) <= REL * ck
        assert (o["cost"] > 0.5).any() and (o["cost"] < 0.5).any()
    finally:
        ctx.close()
