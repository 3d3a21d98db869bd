"""Pins of the oracle's association pieces to third-party code that IS in this image.

The reference searches with pcl::KdTreeFLANN<PointXYZI>::nearestKSearch (GLIO/src/Estimator.cpp:3647, :3746, :3832) =
FLANN's KDTreeSingleIndex over L2_Simple<float>, and fits the plane with Eigen's colPivHouseholderQr (:3661).  Neither
PCL nor Eigen is installed here, but OpenCV's bundled FLANN (cv2.flann_Index; the same FLANN code base: KDTREE_SINGLE =
KDTreeSingleIndex, LINEAR = brute force; its L2<float> functor accumulates a 3-vector as ((dx*dx) + dy*dy) + dz*dz, the
same order as L2_Simple) and LAPACK's column-pivoted QR (scipy.linalg.qr(pivoting=True) = dgeqp3) are.  These tests tie
the oracle's kNN (indices AND float distances, bit for bit) and the 5x3 least-squares solve to those libraries."""
import numpy as np
import pytest

from glio_b200 import synth

cv2 = pytest.importorskip("cv2")
FLANN_INDEX_LINEAR, FLANN_INDEX_KDTREE_SINGLE = 0, 4


def _flann_knn5(map_xyz, qry, algorithm):
    prm = dict(algorithm=algorithm, leaf_max_size=15) if algorithm == FLANN_INDEX_KDTREE_SINGLE else dict(algorithm=algorithm)
    index = cv2.flann_Index(np.ascontiguousarray(map_xyz, np.float32), prm)
    idx, sqd = index.knnSearch(np.ascontiguousarray(qry, np.float32), 5, params=dict(checks=-1, eps=0.0, sorted=True))
    return idx.astype(np.int32), sqd.astype(np.float32)


def test_knn_matches_flann_kdtree_single_and_linear_small(oracle):
    """Small, dense cloud with many near neighbours: FLANN KDTREE_SINGLE == FLANN LINEAR == oracle brute == oracle kd-tree."""
    P = synth.window_problem(W=2, Q=3000, M=40000, seed=31)
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][0, :3], P["poses_init"][0, 3:])
    pm = oracle.transform_points(P["scans"][0], t2, q2)
    ib, db, tie = oracle.knn5_brute(P["map_xyz"], pm)
    assert not tie.any(), "the generator is supposed to be tie free"
    ik, dk = oracle.KdTree(P["map_xyz"]).knn5(pm)
    for algo in (FLANN_INDEX_KDTREE_SINGLE, FLANN_INDEX_LINEAR):
        fi, fd = _flann_knn5(P["map_xyz"], pm, algo)
        assert np.array_equal(fi, ib) and np.array_equal(fd, db), f"oracle brute force differs from FLANN algorithm {algo}"
        assert np.array_equal(fi, ik) and np.array_equal(fd, dk), f"oracle kd-tree differs from FLANN algorithm {algo}"


def test_knn_matches_flann_at_full_map_size(oracle):
    """cfg-2 map (M = 1 M) and 40 k transformed queries of two scans: the oracle's kd-tree (what every full-size parity test
    uses as its reference) returns FLANN KDTreeSingleIndex's indices and float distances bit for bit."""
    P = synth.window_problem(W=20, Q=100_000, M=1_000_000, seed=synth.SEED0 + 2)
    tree = oracle.KdTree(P["map_xyz"])
    index = cv2.flann_Index(P["map_xyz"], dict(algorithm=FLANN_INDEX_KDTREE_SINGLE, leaf_max_size=15))
    for k in (3, 17):
        t2, q2 = synth.lidar_pose_in_world(P["poses_init"][k, :3], P["poses_init"][k, 3:])
        pm = oracle.transform_points(P["scans"][k][:20000], t2, q2)
        ik, dk = tree.knn5(pm)
        fi, fd = index.knnSearch(pm, 5, params=dict(checks=-1, eps=0.0, sorted=True))
        assert np.array_equal(fi.astype(np.int32), ik), "kNN indices differ from FLANN at M = 1M"
        assert np.array_equal(fd.astype(np.float32), dk), "kNN float distances differ from FLANN at M = 1M"


def test_assoc_gate_and_indices_consistent_with_flann(oracle):
    """The association's idx5/sqd5 outputs (what the GPU is compared with) are FLANN's, and the radius gate is applied to
    FLANN's 5th SQUARED distance (quirk Q1, Estimator.cpp:3651)."""
    P = synth.window_problem(W=3, Q=5000, M=60000, seed=77)
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][1, :3], P["poses_init"][1, 3:])
    o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][1], t2, q2)
    fi, fd = _flann_knn5(P["map_xyz"], o["pm"], FLANN_INDEX_KDTREE_SINGLE)
    assert np.array_equal(o["idx5"], fi) and np.array_equal(o["sqd5"], fd)
    assert np.array_equal(o["status"] == oracle.GO_FAIL_RADIUS, ~(fd[:, 4].astype(np.float64) < 1.5))


def test_plane_solve_matches_lapack_pivoted_qr(oracle):
    """colPivHouseholderQr(A).solve(-1) for 5x3 A (Estimator.cpp:3649-3661) against LAPACK dgeqp3 (scipy.linalg.qr with
    pivoting) and against the SVD-based lstsq: <= 1e-12 relative on well-conditioned neighbourhoods, and the same pivot
    choice (largest remaining column norm first)."""
    sl = pytest.importorskip("scipy.linalg")
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(400):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        c = rng.uniform(-60, 60, 3)
        c += n * (3.0 + abs(rng.normal())) * np.sign(n @ c if n @ c != 0 else 1.0)     # keep the plane away from the origin
        basis = np.linalg.svd(n[None, :])[2][1:]
        A = c + rng.uniform(-0.4, 0.4, (5, 2)) @ basis + 0.02 * rng.normal(size=(5, 1)) * n
        A = A.astype(np.float32).astype(np.float64)
        x, rank = oracle.plane_solve5(A)
        assert rank == 3
        Q, R, piv = sl.qr(A, mode="economic", pivoting=True)
        y = sl.solve_triangular(R, Q.T @ (-np.ones(5)))
        x_qr = np.empty(3); x_qr[piv] = y
        x_ls = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        cond = np.linalg.cond(A)
        tol = 1e-13 * cond                       # backward-stable solvers agree to O(eps * cond)
        worst = max(worst, np.max(np.abs(x - x_qr)) / np.max(np.abs(x_qr)) / max(cond, 1.0))
        assert np.max(np.abs(x - x_qr)) <= tol * np.max(np.abs(x_qr))
        assert np.max(np.abs(x - x_ls)) <= 10 * tol * np.max(np.abs(x_ls))
        # first pivot = the column of largest norm (Eigen and LAPACK agree on this rule)
        assert piv[0] == int(np.argmax(np.linalg.norm(A, axis=0)))
    assert worst < 1e-13


def test_plane_unit_normal_matches_lapack_on_synthetic_neighbourhoods(oracle):
    """End to end on real neighbourhoods of the synthetic map: the oracle's unit normal / offset (Estimator.cpp:3662-3663)
    equal the ones computed from LAPACK's pivoted QR solution to 1e-12."""
    sl = pytest.importorskip("scipy.linalg")
    P = synth.window_problem(W=2, Q=2000, M=30000, seed=9)
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][0, :3], P["poses_init"][0, 3:])
    o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][0], t2, q2)
    ok = np.nonzero(o["status"] != oracle.GO_FAIL_RADIUS)[0][:500]
    for i in ok:
        A = P["map_xyz"][o["idx5"][i]].astype(np.float64)
        Q, R, piv = sl.qr(A, mode="economic", pivoting=True)
        y = sl.solve_triangular(R, Q.T @ (-np.ones(5)))
        x = np.empty(3); x[piv] = y
        nrm = np.linalg.norm(x)
        n_ref, d_ref = x / nrm, 1.0 / nrm
        scale = 1e-13 * np.linalg.cond(A) + 1e-12
        assert np.max(np.abs(o["plane"][i, :3] - n_ref)) <= scale and abs(o["plane"][i, 3] - d_ref) <= scale * max(1.0, abs(d_ref))


def test_oracle_lm_dense_qr_equals_lm_normal_equations(oracle):
    """Levenberg-Marquardt in the oracle (ceres.tgz::internal/ceres/levenberg_marquardt_strategy.cc:69-160): the literal DENSE_QR
    linear solve of [J; D] (the front end's options, LidarOdometry.cpp:521-530) and the normal-equation Cholesky solve give
    the same iterates to rounding, the radius follows Ceres' rule, and the path differs from the dogleg one."""
    P = synth.window_problem(W=1, Q=3000, M=40000, seed=synth.SEED0 + 9)
    ident_q, zero_t = [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    t2, q2 = synth.lidar_pose_in_world(P["poses_init"][0, :3], P["poses_init"][0, 3:7])
    state = np.concatenate([t2, q2])[None, :]
    prm = oracle.default_params(); prm.kd_max_radius = 1.0; prm.surf_dist_thres = 0.06; prm.weight_min = 0.4; prm.lidar_const = 1.0
    o = oracle.assoc_scan_to_map(P["map_xyz"], P["scans"][0], t2, q2, prm=prm)
    v = o["status"] == oracle.GO_VALID
    kf = np.zeros(int(v.sum()), np.int32); ones = np.ones(int(v.sum()))
    prob = oracle.WindowProblem(state, None, ident_q, zero_t, huber_delta=0.1)
    prob.add_unary(kf, P["scans"][0][v], o["nsd"][v], ones)
    res = {}
    for strat in (0, 1, 2):
        prob.reset_state(state)
        res[strat] = prob.solve(oracle.solver_options(reserved=strat, max_num_iterations=8), mode=0)
    a, b = res[1], res[2]
    assert a["summary"].num_iterations == b["summary"].num_iterations >= 3
    for x, y in zip(a["steps"], b["steps"]):
        assert np.max(np.abs(x - y)) <= 1e-9
    for ia, ib in zip(a["iterations"], b["iterations"]):
        assert ia["step_is_successful"] == ib["step_is_successful"]
        assert ia["trust_region_radius"] == pytest.approx(ib["trust_region_radius"], rel=1e-8)
    # Ceres' rule on the accepted steps: radius_{k+1} = min(max_radius, radius_k / max(1/3, 1 - (2 rho - 1)^3))
    its = a["iterations"]
    for prev, cur in zip(its, its[1:]):
        if cur["step_is_successful"]:
            want = min(1e16, prev["trust_region_radius"] / max(1.0 / 3.0, 1.0 - (2.0 * cur["relative_decrease"] - 1.0) ** 3))
            assert cur["trust_region_radius"] == pytest.approx(want, rel=1e-12)
    assert a["summary"].final_cost < 0.7 * a["summary"].initial_cost
    # a different algorithm from the dogleg one: the first steps are not the same vector
    assert np.max(np.abs(res[0]["steps"][0] - a["steps"][0])) > 1e-9
