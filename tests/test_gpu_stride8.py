"""The reference's ONLY point layout on this path is pcl::PointXYZI (GLIO/include/utils/common.h:87-89): 32 bytes per
point = x, y, z, padding, intensity, 3 x padding -> stride_floats = 8.  Every entry point that takes points (map, window
scans, single scan, batch frames, local-map clouds) is driven here with that layout — host AND device buffers — and must
give bit for bit what the packed stride-3 call gives and what the oracle gives.  The non-coordinate lanes are filled with
NaN / huge values so that any kernel that reads them poisons its output."""
import numpy as np
import pytest

from glio_b200 import synth

pytestmark = pytest.mark.gpu


def xyzi(xyz, seed=0):
    """(n,3) float32 -> (n,8) float32 PointXYZI rows: [x y z 1(pad) intensity NaN NaN 1e30]."""
    rng = np.random.default_rng(seed)
    out = np.empty((len(xyz), 8), np.float32)
    out[:, :3] = xyz
    out[:, 3] = 1.0                                   # PCL's PCL_ADD_POINT4D puts 1.0f in data[3]
    out[:, 4] = rng.uniform(0, 255, len(xyz))         # intensity
    out[:, 5:7] = np.nan
    out[:, 7] = 1e30
    return out


@pytest.fixture(scope="module")
def ctx():
    from glio_b200 import api
    c = api.Context(0, keep_debug=1)
    yield c
    c.close()


def _assoc_snapshot(ctx, W, Q):
    out = []
    for k in range(W):
        d = ctx.get_assoc_debug(k, Q); m = ctx.get_matches(k, Q)
        out.append((d, m))
    return out


def test_window_path_stride8_equals_stride3_and_oracle(ctx, oracle):
    W, Q, M = 4, 3000, 40000
    P = synth.window_problem(W=W, Q=Q, M=M, seed=5)
    # packed reference run
    ctx.set_map(P["map_xyz"]); ctx.window_set_scans(P["scans"])
    nm3 = ctx.window_associate(P["poses_init"])
    snap3 = _assoc_snapshot(ctx, W, Q)
    blocks3 = ctx.eval_unary(P["poses_init"])
    sol3 = ctx.window_solve(P["poses_init"])
    # PointXYZI run (host buffers)
    ctx.set_map(xyzi(P["map_xyz"], 1)); ctx.window_set_scans([xyzi(s, 2 + k) for k, s in enumerate(P["scans"])])
    nm8 = ctx.window_associate(P["poses_init"])
    snap8 = _assoc_snapshot(ctx, W, Q)
    blocks8 = ctx.eval_unary(P["poses_init"])
    sol8 = ctx.window_solve(P["poses_init"])
    assert np.array_equal(nm3, nm8)
    for (d3, m3), (d8, m8) in zip(snap3, snap8):
        for key in ("pm", "status", "idx5", "sqd5", "plane"):
            assert np.array_equal(d3[key], d8[key]), key
        for key in ("cp", "nsd", "weight", "src"):
            assert np.array_equal(m3[key], m8[key]), key
    for key in ("H", "g", "cost"):
        assert np.array_equal(blocks3[key], blocks8[key]), key
    assert np.array_equal(sol3["poses"], sol8["poses"]) and sol3["summary"].num_iterations == sol8["summary"].num_iterations
    # and against the oracle
    tree = oracle.KdTree(P["map_xyz"])
    for k in range(W):
        t2, q2 = ctx.lidar_pose(P["poses_init"][k])
        o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][k], t2, q2, tree=tree)
        d8, m8 = snap8[k]
        assert np.array_equal(d8["status"], o["status"]) and np.array_equal(d8["pm"], o["pm"])
        ok = o["status"] != oracle.GO_FAIL_RADIUS
        assert np.array_equal(d8["idx5"][ok], o["idx5"][ok]) and np.array_equal(d8["sqd5"][ok], o["sqd5"][ok])
        v = o["status"] == oracle.GO_VALID
        assert np.array_equal(m8["cp"], P["scans"][k][v]) and np.array_equal(m8["nsd"], o["nsd"][v])


def test_single_scan_call_stride8_host_and_device(ctx, oracle):
    torch = pytest.importorskip("torch")
    P = synth.window_problem(W=2, Q=2500, M=30000, seed=17)
    t2, q2 = ctx.lidar_pose(P["poses_init"][1])
    o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][1], t2, q2)
    ok = o["status"] != oracle.GO_FAIL_RADIUS
    m8 = xyzi(P["map_xyz"], 3); s8 = xyzi(P["scans"][1], 4)
    for where in ("host", "device"):
        if where == "host":
            ctx.set_map(m8); n = ctx.assoc_scan_to_map(0, s8, t2, q2)
        else:
            dm = torch.from_numpy(m8).cuda(); ds = torch.from_numpy(s8).cuda()
            ctx.set_map(dm); n = ctx.assoc_scan_to_map(0, ds, t2, q2)
        d = ctx.get_assoc_debug(0, len(s8))
        assert n == o["nvalid"], where
        assert np.array_equal(d["status"], o["status"]) and np.array_equal(d["idx5"][ok], o["idx5"][ok]) and np.array_equal(d["plane"][ok], o["plane"][ok]), where
        m = ctx.get_matches(0, len(s8))
        assert np.array_equal(m["cp"], P["scans"][1][o["status"] == oracle.GO_VALID]), where


def test_batch_frames_stride8(oracle):
    from glio_b200 import api
    K, Q, sr = 4, 3000, 1
    B = synth.batch_problem(K=K, Q=Q, seed=41, search_range=sr, rng_range=5.0)
    cur, oth = [], []
    for i in range(K):
        for j in range(max(0, i - sr), min(K, i + sr + 1)):
            if j != i:
                cur.append(i); oth.append(j)
    res = {}
    for stride in (3, 8):
        c = api.Context(0)
        try:
            for k in range(K):
                c.batch_set_frame(k, B["scans"][k] if stride == 3 else xyzi(B["scans"][k], 10 + k), B["poses_init"][k])
            nm = c.batch_associate_pairs(cur, oth)
            ms = [c.batch_get_matches(int(a), int(b), Q) for a, b in zip(cur, oth)]
            ev = c.eval_binary(B["poses_init"])
            so = c.batch_solve(B["poses_init"], options=api.batch_solver_options(max_num_iterations=4))
            res[stride] = (nm, ms, ev, so)
        finally:
            c.close()
    assert np.array_equal(res[3][0], res[8][0])
    for a, b in zip(res[3][1], res[8][1]):
        for key in ("cp", "weight", "normal_cent", "src"):
            assert np.array_equal(a[key], b[key]), key
    for key in ("Hdiag", "Hoff", "g"):
        assert np.array_equal(res[3][2][key], res[8][2][key]), key
    assert np.array_equal(res[3][3]["poses"], res[8][3]["poses"])
    # one pair against the oracle (local-frame normal / centroid read the STRIDED local points of the searched frame)
    i = 1
    ro = oracle.assoc_pair(B["scans"][cur[i]], B["poses_init"][cur[i], :3], B["poses_init"][cur[i], 3:], B["scans"][oth[i]],
                           B["poses_init"][oth[i], :3], B["poses_init"][oth[i], 3:])
    v = ro["status"] == oracle.GO_VALID
    m = res[8][1][i]
    assert m["n"] == int(v.sum()) and np.array_equal(m["normal_cent"], ro["normal_cent"][v]) and np.array_equal(m["weight"], ro["weight"][v])


def test_localmap_push_stride8(oracle):
    from glio_b200 import api
    rng = np.random.default_rng(3)
    sc = synth.Scene(-60.0, 66.0, rng)
    poses = synth.trajectory(5, rng)
    clouds = [synth.scan_in_lidar_frame(sc, poses[k], 3000, rng) for k in range(5)]
    c = api.Context(0)
    try:
        world = []
        for k, cl in enumerate(clouds):
            t2, q2 = synth.lidar_pose_in_world(poses[k, :3], poses[k, 3:7])
            c.localmap_push(xyzi(cl, k), t2, q2)
            world.append(oracle.transform_points(cl, t2, q2))
        n = c.localmap_build(0.4)
        ref, _ = oracle.voxel_filter(np.concatenate(world), 0.4, stable=True)
        assert n == len(ref) and np.array_equal(c.get_map(), ref)
    finally:
        c.close()
