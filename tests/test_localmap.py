"""Local map maintenance: keyframe clouds -> world frame -> concatenation -> pcl::VoxelGrid(0.4 m).
CPU: the oracle's restatement of PCL's published algorithm against an independent numpy formulation and its own literal
(std::sort) variant.  GPU: the C-ABI path against the oracle (stable order) bit-for-bit, then association on the
device-built map against association on the same map handed over by the host."""
import numpy as np
import pytest

from glio_b200 import synth


def _np_voxel(x, leaf):
    inv = np.float32(1) / np.float32(leaf)
    mb = np.floor(x.min(0) * inv).astype(np.int64); xb = np.floor(x.max(0) * inv).astype(np.int64); div = xb - mb + 1
    ijk = (np.floor(x * inv) - mb.astype(np.float32)).astype(np.int64)
    vid = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    u, inv_i, cnt = np.unique(vid, return_inverse=True, return_counts=True)
    mean = np.zeros((len(u), 3)); np.add.at(mean, inv_i, x.astype(np.float64)); mean /= cnt[:, None]
    return u, mean, cnt


def _frames(seed=3, K=6, Q=4000):
    rng = np.random.default_rng(seed)
    sc = synth.Scene(-60.0, K + 60.0, rng)
    poses = synth.trajectory(K, rng)
    clouds = [synth.scan_in_lidar_frame(sc, poses[k], Q, rng) for k in range(K)]
    return clouds, poses


def test_oracle_voxel_filter_against_numpy(oracle):
    rng = np.random.default_rng(11)
    x = (rng.random((30000, 3)) * [40, 20, 4] - [7, 9, 1.5]).astype(np.float32)
    for leaf in (0.4, 0.2, 1.0):
        a, ia = oracle.voxel_filter(x, leaf, stable=True)
        b, ib = oracle.voxel_filter(x, leaf, stable=False)
        u, mean, cnt = _np_voxel(x, leaf)
        assert np.array_equal(ia, u) and np.array_equal(ib, u)           # same voxels, ascending voxel index
        assert np.abs(a - mean).max() < 2e-5 and np.abs(b - mean).max() < 2e-5   # float sums of <= ~60 points at |x| < 40
        assert np.abs(a - b).max() < 2e-5                                 # the within-voxel order only moves the last bits
    # a single point per voxel comes back unchanged; PCL's overflow guard passes the input through
    y = (np.arange(30, dtype=np.float32)[:, None] * np.array([[1.0, 2.0, 3.0]], np.float32))
    a, _ = oracle.voxel_filter(y, 0.4)
    assert np.array_equal(np.sort(a, axis=0), np.sort(y, axis=0))
    z = np.array([[0, 0, 0], [4000, 4000, 4000]], np.float32)
    assert oracle.voxel_filter(z, 0.001) is None


@pytest.mark.gpu
def test_device_localmap_matches_oracle(oracle):
    from glio_b200 import api
    clouds, poses = _frames()
    ctx = api.Context(0, keep_debug=1)
    try:
        world = []
        for k, cl in enumerate(clouds):
            t2, q2 = synth.lidar_pose_in_world(poses[k, :3], poses[k, 3:7])      # q_po*q_bl, q_po*t_bl + t_po (Estimator.cpp:3563-3564)
            ctx.localmap_push(cl, t2, q2, at_front=False)
            world.append(oracle.transform_points(cl, t2, q2))
        assert ctx.localmap_size() == (len(clouds), sum(len(c) for c in clouds))
        n = ctx.localmap_build(0.4)
        ref, vid = oracle.voxel_filter(np.concatenate(world), 0.4, stable=True)
        got = ctx.get_map()
        assert n == len(ref) == len(got)
        assert np.array_equal(got, ref), "voxel-filtered local map differs (bit-exact expected: same float sums in the same order)"
        # deque semantics: pop the oldest, push a new one at the back, one at the front
        ctx.localmap_pop_front(); world.pop(0)
        t2, q2 = synth.lidar_pose_in_world(poses[0, :3], poses[0, 3:7])
        ctx.localmap_push(clouds[0], t2, q2, at_front=True); world.insert(0, oracle.transform_points(clouds[0], t2, q2))
        n2 = ctx.localmap_build(0.4)
        ref2, _ = oracle.voxel_filter(np.concatenate(world), 0.4, stable=True)
        assert n2 == len(ref2) and np.array_equal(ctx.get_map(), ref2)
        # leaf <= 0: the concatenation itself is the map
        n3 = ctx.localmap_build(0.0)
        assert n3 == sum(len(w) for w in world) and np.array_equal(ctx.get_map(), np.concatenate(world))
        # association on the device-built map == association on the same points set from the host
        ctx.localmap_build(0.4)
        t2, q2 = synth.lidar_pose_in_world(poses[-1, :3] + 0.03, poses[-1, 3:7])
        na = ctx.assoc_scan_to_map(0, clouds[-1], t2, q2)
        da = ctx.get_assoc_debug(0, len(clouds[-1]))
        ctx.set_map(ref2)
        nb = ctx.assoc_scan_to_map(0, clouds[-1], t2, q2)
        db = ctx.get_assoc_debug(0, len(clouds[-1]))
        assert na == nb and na > 100
        for k in ("status", "idx5", "sqd5", "plane"):
            assert np.array_equal(da[k], db[k]), k
        o = oracle.assoc_scan_to_map(ref2, clouds[-1], t2, q2)
        assert np.array_equal(da["status"], o["status"])
        ctx.localmap_clear()
        assert ctx.localmap_size() == (0, 0)
    finally:
        ctx.close()
