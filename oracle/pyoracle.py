"""ctypes binding of the CPU oracle (oracle/libglio_oracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under glio_b200/ may import this module.
PARITY UNPINNED: see oracle/glio_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GO_VALID, GO_FAIL_RADIUS, GO_FAIL_PLANE, GO_FAIL_WEIGHT = 0, 1, 2, 3


class AssocParams(C.Structure):
    _fields_ = [("kd_max_radius", C.c_double), ("surf_dist_thres", C.c_double),
                ("lidar_const", C.c_double), ("weight_min", C.c_double),
                ("batch_max_radius", C.c_double), ("batch_dist_thres", C.c_double),
                ("batch_score", C.c_double)]


def default_params():
    # GLIO/config/config_urban_hk.yaml:70-72, Estimator.cpp:3681,3751,3778,3798
    return AssocParams(1.5, 0.18, 7.5, 0.3, 1.5, 0.18, 2.5)


def build(force=False):
    so = os.path.join(_HERE, "libglio_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libglio_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.go_kdtree_build.restype = C.c_void_p
        _LIB.go_assoc_scan_to_map.restype = C.c_int64
        _LIB.go_assoc_pair.restype = C.c_int64
        _LIB.go_plane_solve5.restype = C.c_int
        if hasattr(_LIB, "go_solver_create"):
            _LIB.go_solver_create.restype = C.c_void_p
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def transform_points(xyz, t, q):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(xyz)
    t = _f64(t); q = _f64(q)
    lib().go_transform_points(_p(xyz), C.c_int64(len(xyz)), _p(t), _p(q), _p(out))
    return out


def voxel_filter(xyz, leaf=0.4, stable=True):
    """pcl::VoxelGrid restatement (see glio_oracle.h).  Returns (points, voxel_idx) or None when PCL would pass the input through."""
    x = _f32(xyz).reshape(-1, 3); n = len(x)
    out = np.empty((n, 3), np.float32); idx = np.empty(n, np.int32)
    lib().go_voxel_filter.restype = C.c_int64
    m = lib().go_voxel_filter(_p(x), C.c_int64(n), C.c_float(leaf), C.c_int(1 if stable else 0), _p(out), _p(idx))
    if m < 0:
        return None
    return out[:m].copy(), idx[:m].copy()


def knn5_brute(map_xyz, qry_xyz):
    m = _f32(map_xyz).reshape(-1, 3); q = _f32(qry_xyz).reshape(-1, 3)
    idx = np.empty((len(q), 5), np.int32); sqd = np.empty((len(q), 5), np.float32)
    tie = np.empty(len(q), np.uint8)
    lib().go_knn5_brute(_p(m), C.c_int64(len(m)), _p(q), C.c_int64(len(q)), _p(idx), _p(sqd), _p(tie))
    return idx, sqd, tie


def extract_features(cloud_xyzi, scan_start, scan_end, ds_rate=1, edge_thres=1.0, surf_thres=0.1, ds_v=0.4, stable=True):
    """Preprocessing::cloudHandler's feature extraction (GLIO/src/Preprocessing.cpp:529-655); see glio_oracle_features.cpp."""
    cl = _f32(cloud_xyzi).reshape(-1, 4); n = len(cl)
    ss = np.ascontiguousarray(scan_start, np.int32); se = np.ascontiguousarray(scan_end, np.int32); S = len(ss)
    out = dict(curvature=np.empty(n, np.float32), label=np.empty(n, np.int8), sharp=np.empty(12 * S, np.int32), less_sharp=np.empty(60 * S, np.int32),
               flat=np.empty(24 * S, np.int32), less_flat=np.empty(n, np.int32), less_flat_ds=np.empty((n, 4), np.float32), ring_ds_count=np.empty(S, np.int32))
    cnt = [C.c_int64(0) for _ in range(5)]
    rc = lib().go_extract_features(_p(cl), C.c_int64(n), C.c_int(S), _p(ss), _p(se), C.c_int(ds_rate), C.c_double(edge_thres), C.c_double(surf_thres),
                                   C.c_float(ds_v), C.c_int(1 if stable else 0), _p(out["curvature"]), _p(out["label"]), _p(out["sharp"]), C.byref(cnt[0]),
                                   _p(out["less_sharp"]), C.byref(cnt[1]), _p(out["flat"]), C.byref(cnt[2]), _p(out["less_flat"]), C.byref(cnt[3]),
                                   _p(out["less_flat_ds"]), C.byref(cnt[4]), _p(out["ring_ds_count"]))
    assert rc == 0
    for k, c in zip(("sharp", "less_sharp", "flat", "less_flat", "less_flat_ds"), cnt):
        out[k] = out[k][:c.value]
    return out


class KdTree:
    def __init__(self, xyz):
        self.xyz = _f32(xyz).reshape(-1, 3)
        self.h = C.c_void_p(lib().go_kdtree_build(_p(self.xyz), C.c_int64(len(self.xyz))))

    def knn5(self, qry):
        q = _f32(qry).reshape(-1, 3)
        idx = np.empty((len(q), 5), np.int32); sqd = np.empty((len(q), 5), np.float32)
        lib().go_kdtree_knn5(self.h, _p(q), C.c_int64(len(q)), _p(idx), _p(sqd))
        return idx, sqd

    def __del__(self):
        try:
            if self.h:
                lib().go_kdtree_free(self.h)
                self.h = None
        except Exception:
            pass


def plane_solve5(A):
    A = _f64(A).reshape(5, 3)
    x = np.zeros(3)
    npiv = lib().go_plane_solve5(_p(A), _p(x))
    return x, npiv


def assoc_scan_to_map(map_xyz, scan_xyz, t, q, prm=None, tree=None, nthreads=0):
    prm = prm or default_params()
    m = _f32(map_xyz).reshape(-1, 3); s = _f32(scan_xyz).reshape(-1, 3)
    Q = len(s)
    out = dict(status=np.empty(Q, np.uint8), idx5=np.empty((Q, 5), np.int32), sqd5=np.empty((Q, 5), np.float32),
               pm=np.empty((Q, 3), np.float32), plane=np.empty((Q, 4), np.float64),
               nsd=np.empty((Q, 4), np.float32), weight=np.empty(Q, np.float32), score=np.empty(Q, np.float64))
    t = _f64(t); q = _f64(q)
    n = lib().go_assoc_scan_to_map(C.byref(prm), _p(m), C.c_int64(len(m)), tree.h if tree is not None else None,
                                   _p(s), C.c_int64(Q), _p(t), _p(q), _p(out["status"]), _p(out["idx5"]),
                                   _p(out["sqd5"]), _p(out["pm"]), _p(out["plane"]), _p(out["nsd"]),
                                   _p(out["weight"]), _p(out["score"]), C.c_int(nthreads))
    out["nvalid"] = int(n)
    return out


def assoc_pair(cur_xyz, t_c, q_c, oth_xyz, t_o, q_o, prm=None, use_kdtree=True, nthreads=0):
    prm = prm or default_params()
    c = _f32(cur_xyz).reshape(-1, 3); o = _f32(oth_xyz).reshape(-1, 3)
    Q = len(c)
    out = dict(status=np.empty(Q, np.uint8), idx5=np.empty((Q, 5), np.int32), sqd5=np.empty((Q, 5), np.float32),
               weight=np.empty(Q, np.float32), score=np.empty(Q, np.float64), normal_cent=np.empty((Q, 6), np.float64))
    n = lib().go_assoc_pair(C.byref(prm), _p(c), C.c_int64(Q), _p(_f64(t_c)), _p(_f64(q_c)), _p(o),
                            C.c_int64(len(o)), _p(_f64(t_o)), _p(_f64(q_o)), C.c_int(1 if use_kdtree else 0),
                            _p(out["status"]), _p(out["idx5"]), _p(out["sqd5"]), _p(out["weight"]),
                            _p(out["score"]), _p(out["normal_cent"]), C.c_int(nthreads))
    out["nvalid"] = int(n)
    return out


def eval_unary(poses, q_lb, t_lb, kf, cp, nsd, score, huber_delta=1.0, mode=0, jac_kind=0, per_residual=True):
    poses = _f64(poses).reshape(-1, 7); W = len(poses)
    kf = np.ascontiguousarray(kf, np.int32); cp = _f32(cp).reshape(-1, 3); nsd = _f32(nsd).reshape(-1, 4)
    score = _f64(score); N = len(kf)
    r = np.empty(N) if per_residual else None
    J = np.empty((N, 6)) if per_residual else None
    c = np.empty(N) if per_residual else None
    H = np.zeros((6 * W, 6 * W)); g = np.zeros(6 * W); ct = np.zeros(1)
    lib().go_eval_unary(C.c_int(mode), C.c_int(jac_kind), C.c_int(W), _p(poses), _p(_f64(q_lb)), _p(_f64(t_lb)),
                        C.c_double(huber_delta), C.c_int64(N), _p(kf), _p(cp), _p(nsd), _p(score),
                        _p(r), _p(J), _p(c), _p(H), _p(g), _p(ct))
    return dict(r=r, J=J, cost=c, H=H, g=g, cost_total=float(ct[0]))


def eval_binary(poses, kf_c, kf_o, cp, normal_cent, score, huber_delta=0.0, mode=0, per_residual=True):
    poses = _f64(poses).reshape(-1, 7); K = len(poses)
    kf_c = np.ascontiguousarray(kf_c, np.int32); kf_o = np.ascontiguousarray(kf_o, np.int32)
    cp = _f32(cp).reshape(-1, 3); nc = _f64(normal_cent).reshape(-1, 6); score = _f64(score); N = len(kf_c)
    r = np.empty(N) if per_residual else None
    J = np.empty((N, 12)) if per_residual else None
    c = np.empty(N) if per_residual else None
    H = np.zeros((6 * K, 6 * K)); g = np.zeros(6 * K); ct = np.zeros(1)
    lib().go_eval_binary(C.c_int(mode), C.c_int(K), _p(poses), C.c_double(huber_delta), C.c_int64(N), _p(kf_c),
                         _p(kf_o), _p(cp), _p(nc), _p(score), _p(r), _p(J), _p(c), _p(H), _p(g), _p(ct))
    return dict(r=r, J=J, cost=c, H=H, g=g, cost_total=float(ct[0]))


def eval_edge(poses, q_lb, t_lb, kf, cp, pa, pb, s, huber_delta=1.0, mode=0, per_residual=True):
    poses = _f64(poses).reshape(-1, 7); W = len(poses)
    kf = np.ascontiguousarray(kf, np.int32); cp = _f32(cp).reshape(-1, 3)
    pa = _f32(pa).reshape(-1, 3); pb = _f32(pb).reshape(-1, 3); s = _f64(s); N = len(kf)
    r = np.empty(N) if per_residual else None
    J = np.empty((N, 6)) if per_residual else None
    c = np.empty(N) if per_residual else None
    H = np.zeros((6 * W, 6 * W)); g = np.zeros(6 * W); ct = np.zeros(1)
    lib().go_eval_edge(C.c_int(mode), C.c_int(W), _p(poses), _p(_f64(q_lb)), _p(_f64(t_lb)), C.c_double(huber_delta),
                       C.c_int64(N), _p(kf), _p(cp), _p(pa), _p(pb), _p(s), _p(r), _p(J), _p(c), _p(H), _p(g), _p(ct))
    return dict(r=r, J=J, cost=c, H=H, g=g, cost_total=float(ct[0]))


def huber(a, s):
    rho = np.zeros(3); lib().go_huber(C.c_double(a), C.c_double(s), _p(rho)); return rho


def corrector(sq_norm, rho):
    out = np.zeros(3); lib().go_corrector(C.c_double(sq_norm), _p(_f64(rho)), _p(out)); return out


def quat_plus(x, delta):
    out = np.zeros(4); lib().go_quat_plus(_p(_f64(x)), _p(_f64(delta)), _p(out)); return out


def quat_plus_jacobian(x):
    out = np.zeros(12); lib().go_quat_plus_jacobian(_p(_f64(x)), _p(out)); return out.reshape(4, 3)


# ---------------------------------------------------------------------------------------------------
# window problem + Ceres-semantics solve (oracle/glio_oracle_solver.cpp)
# ---------------------------------------------------------------------------------------------------
class OSolverOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("dogleg_type", C.c_int32), ("use_nonmonotonic_steps", C.c_int32),
                ("max_consecutive_nonmonotonic_steps", C.c_int32), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("fuse_candidate_jacobian", C.c_int32), ("reserved", C.c_int32)]


class OIteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32), ("reserved", C.c_int32),
                ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double),
                ("step_norm", C.c_double), ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double), ("mu", C.c_double)]


class OSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("num_evaluations", C.c_int32), ("num_jacobian_evaluations", C.c_int32),
                ("num_linear_solves", C.c_int32), ("num_valid_steps", C.c_int32), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("message", C.c_char * 128)]


def solver_options(**kw):
    # ceres.tgz::include/ceres/solver.h defaults + Estimator.cpp:2424-2433
    o = OSolverOptions(15, 0, 0, 5, 1e4, 1e16, 1e-32, 1e-3, 1e-6, 1e32, 5, 1, 1e-6, 1e-10, 1e-8, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class WindowProblem:
    def __init__(self, poses, speed_bias, q_lb, t_lb, huber_delta=1.0):
        L = lib()
        L.go_problem_create.restype = C.c_void_p
        self.W = len(np.asarray(poses).reshape(-1, 7))
        self.has_sb = speed_bias is not None
        pb = _f64(poses).reshape(-1, 7); sb = None if speed_bias is None else _f64(speed_bias).reshape(-1, 9)
        self.h = C.c_void_p(L.go_problem_create(C.c_int(self.W), _p(pb), _p(sb), _p(_f64(q_lb)), _p(_f64(t_lb)), C.c_double(huber_delta)))
        self.n = self.W * (15 if self.has_sb else 6)

    def add_unary(self, kf, cp, nsd, score):
        kf = np.ascontiguousarray(kf, np.int32); cp = _f32(cp).reshape(-1, 3); nsd = _f32(nsd).reshape(-1, 4); score = _f64(score)
        lib().go_problem_add_unary(self.h, C.c_int64(len(kf)), _p(kf), _p(cp), _p(nsd), _p(score))

    def add_binary(self, kf_c, kf_o, cp, normal_cent, score):
        kf_c = np.ascontiguousarray(kf_c, np.int32); kf_o = np.ascontiguousarray(kf_o, np.int32)
        cp = _f32(cp).reshape(-1, 3); nc = _f64(normal_cent).reshape(-1, 6); score = _f64(score)
        lib().go_problem_add_binary(self.h, C.c_int64(len(kf_c)), _p(kf_c), _p(kf_o), _p(cp), _p(nc), _p(score))

    def add_prior(self, kf, t0, q0, sb0, sqrt_w):
        lib().go_problem_add_prior(self.h, C.c_int(kf), _p(_f64(t0)), _p(_f64(q0)), _p(None if sb0 is None else _f64(sb0)), _p(_f64(sqrt_w)))

    def add_between(self, i, j, dp, dq, dv, dt, sqrt_w):
        lib().go_problem_add_between(self.h, C.c_int(i), C.c_int(j), _p(_f64(dp)), _p(_f64(dq)), _p(_f64(dv)), C.c_double(dt), _p(_f64(sqrt_w)))

    def add_range(self, kf, lever, sat, rho, w):
        lib().go_problem_add_range(self.h, C.c_int(kf), _p(_f64(lever)), _p(_f64(sat)), C.c_double(rho), C.c_double(w))

    def set_marg_prior(self, prior):
        """prior: dict(W, lin_jac, lin_res, x0_pose, x0_sb) in the numbering of THIS window (None removes it)."""
        if prior is None:
            lib().go_problem_set_marg_prior(self.h, C.c_int(0), None, None, None, None)
            return
        a = [_f64(prior[k]) for k in ("lin_jac", "lin_res", "x0_pose", "x0_sb")]
        lib().go_problem_set_marg_prior(self.h, C.c_int(int(prior["W"])), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]))

    def marginalize(self, eps=1e-8, mode=0):
        """MarginalizationInfo::PreMarginalize + Marginalize at the problem's current state (GLIO/src/MarginalizationFactor.cpp:107-202)."""
        W = self.W; N = 6 * W + 18; n = N - 15
        out = dict(W=W, lin_jac=np.zeros((n, n)), lin_res=np.zeros(n), x0_pose=np.zeros((W - 1, 7)), x0_sb=np.zeros(9), A=np.zeros((N, N)), b=np.zeros(N))
        rc = lib().go_problem_marginalize(self.h, C.c_double(eps), C.c_int(mode), _p(out["lin_jac"]), _p(out["lin_res"]), _p(out["x0_pose"]), _p(out["x0_sb"]), _p(out["A"]), _p(out["b"]))
        assert rc == 0, rc
        return out

    def reset_state(self, poses, speed_bias=None):
        pb = _f64(poses).reshape(-1, 7); sb = None if speed_bias is None else _f64(speed_bias).reshape(-1, 9)
        lib().go_problem_set_state(self.h, _p(pb), _p(sb))

    def host_normal_eq(self):
        H = np.zeros((self.n, self.n)); g = np.zeros(self.n); c = np.zeros(1)
        lib().go_problem_host_normal_eq(self.h, _p(H), _p(g), _p(c))
        return H, g, float(c[0])

    def solve(self, options=None, mode=0, nthreads=1, max_log=256):
        opt = options or solver_options()
        summ = OSummary(); log = (OIteration * max_log)(); steps = np.zeros((max_log, self.n))
        lib().go_problem_solve(self.h, C.byref(opt), C.c_int(mode), C.c_int(nthreads), C.byref(summ), log, C.c_int(max_log), _p(steps), C.c_int64(steps.size))
        poses = np.zeros((self.W, 7)); sb = np.zeros((self.W, 9)) if self.has_sb else None
        lib().go_problem_get_state(self.h, _p(poses), _p(sb))
        names = [f[0] for f in OIteration._fields_ if f[0] != "reserved"]
        its = [{k: getattr(log[i], k) for k in names} for i in range(min(summ.num_iterations, max_log))]
        return dict(poses=poses, speed_bias=sb, summary=summ, iterations=its, steps=steps[:summ.num_valid_steps])

    def __del__(self):
        try:
            if self.h:
                lib().go_problem_free(self.h); self.h = None
        except Exception:
            pass
