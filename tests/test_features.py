"""the front end's feature extraction (Preprocessing::cloudHandler, GLIO/src/Preprocessing.cpp:529-655).
CPU: the oracle's restatement against an independent numpy formulation of the curvature and of the selection invariants, and
its literal std::sort variant against the tie-by-index one.  GPU: the device pass against the oracle, bit for bit."""
import numpy as np
import pytest

from glio_b200 import synth


@pytest.fixture(scope="module")
def sweep():
    return synth.ring_scan(n_rings=32, n_az=1500)


def test_oracle_feature_extraction_invariants(oracle, sweep):
    cloud, ss, se = sweep
    o = oracle.extract_features(cloud, ss, se, ds_rate=1, edge_thres=1.0, surf_thres=0.1, ds_v=0.4)
    n = len(cloud)
    # curvature: float32, the reference's left-to-right sum (numpy float32 arithmetic in the same order)
    x = cloud[:, :3]
    acc = x[0:n - 10].copy()
    for k in range(1, 5):
        acc = acc + x[k:n - 10 + k]
    acc = acc - np.float32(10) * x[5:n - 5]
    for k in range(6, 11):
        acc = acc + x[k:n - 10 + k]
    curv = (acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1]) + acc[:, 2] * acc[:, 2]
    assert np.array_equal(o["curvature"][5:n - 5], curv.astype(np.float32))
    lab = o["label"]
    assert set(np.unique(lab)).issubset({-1, 0, 1, 2})
    assert np.array_equal(np.sort(o["sharp"]), np.nonzero(lab == 2)[0]) and np.array_equal(np.sort(o["flat"]), np.nonzero(lab == -1)[0])
    assert np.array_equal(np.sort(o["less_sharp"]), np.nonzero(lab >= 1)[0])
    assert len(o["sharp"]) <= 12 * len(ss) and len(o["flat"]) <= 24 * len(ss) and len(o["sharp"]) > 20 and len(o["flat"]) > 200
    assert (o["curvature"][o["less_sharp"]] > 1.0).all() and (o["curvature"][o["flat"]] < 0.1).all()
    # less flat = every point of the processed ranges with label <= 0 (none is within 0.5 m here), in index order
    inside = np.zeros(n, bool)
    for a, b in zip(ss, se):
        if b - a >= 6:
            inside[a:a + (b - a) * 6 // 6] = True          # sectors tile [start, start + (end-start)) : ep of sector 5 = end - 1
    assert np.array_equal(o["less_flat"], np.nonzero(inside & (lab <= 0))[0])
    # the voxel filter shrinks every ring and keeps ring order (intensity = ring + 0.1 relTime averages stay within the ring)
    assert 0 < len(o["less_flat_ds"]) < len(o["less_flat"]) and o["ring_ds_count"].sum() == len(o["less_flat_ds"])
    ring_of = np.floor(o["less_flat_ds"][:, 3] + 1e-4).astype(int)
    assert (np.diff(ring_of) >= 0).all()
    # literal std::sort variant: same labels and sets on this tie-free sweep
    o2 = oracle.extract_features(cloud, ss, se, stable=False)
    assert np.array_equal(o2["label"], lab) and np.array_equal(o2["sharp"], o["sharp"]) and np.array_equal(o2["flat"], o["flat"])
    assert np.abs(o2["less_flat_ds"] - o["less_flat_ds"]).max() < 1e-4
    # ds_rate skips rings
    o3 = oracle.extract_features(cloud, ss, se, ds_rate=2)
    odd = np.zeros(n, bool)
    for r in range(1, len(ss), 2):
        odd[ss[r] - 5:se[r] + 6] = True
    assert (o3["label"][odd] == 0).all() and len(o3["sharp"]) < len(o["sharp"])


@pytest.mark.gpu
@pytest.mark.parametrize("stride,ds_rate", [(4, 1), (8, 2)])
def test_device_feature_extraction_matches_oracle(oracle, sweep, stride, ds_rate):
    from glio_b200 import api
    cloud, ss, se = sweep
    o = oracle.extract_features(cloud, ss, se, ds_rate=ds_rate, edge_thres=1.0, surf_thres=0.1, ds_v=0.4, stable=True)
    buf = cloud
    if stride == 8:                                  # pcl::PointXYZI (common.h:87-89): x y z 1 | intensity pad pad pad
        buf = np.full((len(cloud), 8), np.nan, np.float32)
        buf[:, :3] = cloud[:, :3]; buf[:, 3] = 1.0; buf[:, 4] = cloud[:, 3]
    ctx = api.Context(0)
    try:
        d = ctx.extract_features(buf, ss, se, ds_rate=ds_rate, edge_thres=1.0, surf_thres=0.1, ds_v=0.4)
        assert np.array_equal(d["curvature"], o["curvature"])
        assert np.array_equal(d["label"], o["label"])
        for k in ("sharp", "less_sharp", "flat", "less_flat"):
            assert np.array_equal(d[k], o[k]), k
        assert d["less_flat_ds"].shape == o["less_flat_ds"].shape and np.array_equal(d["less_flat_ds"], o["less_flat_ds"])
        assert ctx.launch_count >= 5
    finally:
        ctx.close()
