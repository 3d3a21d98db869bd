"""Python mirror of the C ABI (include/glio_b200.h) — a thin ctypes layer over glio_b200/libglio_b200.so.

The method names follow the reference's own vocabulary (GLIO/src/Estimator.cpp): set_map ~ kd_tree->setInputCloud,
assoc_scan_to_map ~ findCorrespondingSurfFeatures, select ~ featureSelection (as an input index list),
eval_unary ~ ResidualBlock::Evaluate over LidarPlaneNormFactor + normal-equation accumulation.

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, construction raises.
torch is used only for device buffers/streams by callers (bench.py); this module needs numpy + ctypes only.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libglio_b200.so")
_LIB = None

HOST, DEVICE = 0, 1
MATCH_VALID, MATCH_FAIL_RADIUS, MATCH_FAIL_PLANE, MATCH_FAIL_WEIGHT = 0, 1, 2, 3


class GlioError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("kd_max_radius", C.c_double), ("surf_dist_thres", C.c_double), ("lidar_const", C.c_double),
                ("weight_min", C.c_double), ("huber_delta", C.c_double), ("q_lb", C.c_double * 4),
                ("t_lb", C.c_double * 3), ("batch_max_radius", C.c_double), ("batch_dist_thres", C.c_double),
                ("batch_score", C.c_double), ("cell_size", C.c_float), ("keep_debug", C.c_int32), ("unit_score", C.c_int32)]


def lib():
    """Load the CUDA library; raise loudly when it is not built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIB_PATH):
            raise GlioError(f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(glio_b200 has no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.glio_last_error.restype = C.c_char_p
        L.glio_last_error.argtypes = [C.c_void_p]
        L.glio_stream.restype = C.c_void_p
        L.glio_stream.argtypes = [C.c_void_p]
        L.glio_launch_count.restype = C.c_int64
        L.glio_launch_count.argtypes = [C.c_void_p]
        L.glio_destroy.argtypes = [C.c_void_p]
        L.glio_destroy.restype = None
        L.glio_marg_prior_destroy.argtypes = [C.c_void_p]
        L.glio_marg_prior_destroy.restype = None
        L.glio_marg_prior_create.restype = C.c_void_p
        _LIB = L
    return _LIB


def default_params(**kw):
    p = Params()
    lib().glio_default_params(C.byref(p))
    for k, v in kw.items():
        if k in ("q_lb", "t_lb"):
            arr = getattr(p, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(p, k, v)
    return p


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    raise TypeError(type(a))


def _points_arg(xyz):
    """numpy float32 (n,3|stride) host array, or (device_ptr:int, n, stride) tuple / torch tensor."""
    if isinstance(xyz, np.ndarray):
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(-1, 3)
        return a, _ptr(a), a.shape[0], a.shape[1], HOST
    if hasattr(xyz, "data_ptr"):          # torch tensor on the GPU
        assert xyz.is_cuda and xyz.is_contiguous() and xyz.dtype.itemsize == 4
        return xyz, C.c_void_p(xyz.data_ptr()), xyz.shape[0], xyz.shape[1], DEVICE
    raise TypeError("points must be a numpy array or a CUDA torch tensor")


class Context:
    def __init__(self, device=0, params=None, **kw):
        self._lib = lib()
        self.params = params if params is not None else default_params(**kw)
        self._h = C.c_void_p()
        rc = self._lib.glio_create(C.c_int(device), C.byref(self.params), C.byref(self._h))
        if rc != 0:
            msg = self._lib.glio_last_error(None)
            raise GlioError(f"glio_create failed ({rc}): {msg.decode() if msg else ''}")
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.glio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            msg = self._lib.glio_last_error(self._h)
            raise GlioError(f"glio error {rc}: {msg.decode() if msg else ''}")

    @property
    def stream(self):
        return self._lib.glio_stream(self._h)

    @property
    def launch_count(self):
        return int(self._lib.glio_launch_count(self._h))

    def synchronize(self):
        self._chk(self._lib.glio_synchronize(self._h))

    def lidar_pose(self, pose_body):
        pb = np.ascontiguousarray(pose_body, np.float64)
        t2 = np.zeros(3); q2 = np.zeros(4)
        self._lib.glio_lidar_pose(C.byref(self.params), _ptr(pb), _ptr(t2), _ptr(q2))
        return t2, q2

    KERNELS = ["k_feat_curvature", "k_feat_select", "k_feat_voxel", "k_feat_offsets", "k_feat_gather", "k_knn_tile", "k_knn_tile2", "k_defer_scatter", "k_knn_box_start", "k_knn_grow", "k_knn_box_far", "k_knn_far", "k_knn_box_cells", "k_knn_team", "k_make_pairs", "k_knn_search", "k_knn_deferred", "k_knn_thread", "k_knn_box", "k_plane_fit", "k_plane_fit_pair", "k_eval_unary", "k_eval_unary_cost", "k_transform_hist", "k_order_scatter",
               "k_compact", "k_flags", "k_cell_hist", "k_cell_scatter", "k_load_bounds", "k_scan_block", "k_scan_add",
               "k_eval_binary", "k_eval_binary_cost", "k_bin_assemble",
               "k_lm_transform", "k_vox_hist", "k_vox_scatter", "k_vox_sort", "k_vox_flags", "k_vox_centroid", "k_init_bounds"]

    def lib_profile(self, on):
        self._chk(self._lib.glio_profile_enable(self._h, C.c_int(1 if on else 0)))

    def lib_profile_read(self):
        out = {}
        for k in self.KERNELS:
            ms = C.c_double(0); n = C.c_int64(0)
            self._chk(self._lib.glio_profile_get(self._h, k.encode(), C.byref(ms), C.byref(n)))
            if n.value:
                out[k] = (ms.value, int(n.value))
        return out

    def knn_fallback_queries(self, reset=False):
        n = C.c_int64(0)
        self._chk(self._lib.glio_get_stats(self._h, C.byref(n), C.c_int(1 if reset else 0)))
        return int(n.value)

    def get_match_counts(self, W, active=True):
        nm = np.zeros(W, np.int64); na = np.zeros(W, np.int64)
        self._chk(self._lib.glio_get_match_counts(self._h, C.c_int(W), _ptr(nm), _ptr(na)))
        return na if active else nm

    # ---- K0
    def set_map(self, xyz):
        keep, p, n, stride, mem = _points_arg(xyz)
        self._map_keep = keep
        self._chk(self._lib.glio_set_map(self._h, p, C.c_int64(n), C.c_int(stride), C.c_int(mem)))

    def map_prefetch(self, xyz):
        keep, p, n, stride, mem = _points_arg(xyz)
        assert mem == HOST
        self._map_pf_keep = keep
        self._chk(self._lib.glio_map_prefetch(self._h, p, C.c_int64(n), C.c_int(stride)))

    # ---- K1
    # ---- local map maintenance on the device
    def localmap_clear(self):
        self._chk(self._lib.glio_localmap_clear(self._h))

    def localmap_push(self, cloud_xyz, t, q, at_front=False):
        keep, p, n, stride, mem = _points_arg(cloud_xyz)
        t = np.ascontiguousarray(t, np.float64); q = np.ascontiguousarray(q, np.float64)
        self._chk(self._lib.glio_localmap_push(self._h, C.c_int(1 if at_front else 0), p, C.c_int64(n), C.c_int(stride), C.c_int(mem), _ptr(t), _ptr(q)))

    def localmap_pop_front(self):
        self._chk(self._lib.glio_localmap_pop_front(self._h))

    def localmap_size(self):
        nf = C.c_int(0); npnt = C.c_int64(0)
        self._chk(self._lib.glio_localmap_size(self._h, C.byref(nf), C.byref(npnt)))
        return nf.value, int(npnt.value)

    def localmap_build(self, leaf=0.4):
        n = C.c_int64(0)
        self._chk(self._lib.glio_localmap_build(self._h, C.c_float(leaf), C.byref(n)))
        return int(n.value)

    def get_map(self):
        n = C.c_int64(0)
        self._chk(self._lib.glio_get_map(self._h, C.c_int64(0), None, C.byref(n)))
        out = np.empty((int(n.value), 3), np.float32)
        self._chk(self._lib.glio_get_map(self._h, C.c_int64(len(out)), _ptr(out), C.byref(n)))
        return out

    def assoc_scan_to_map(self, slot, scan_xyz, t, q):
        keep, p, n, stride, mem = _points_arg(scan_xyz)
        self._keep.append(keep); self._keep = self._keep[-64:]
        t = np.ascontiguousarray(t, np.float64); q = np.ascontiguousarray(q, np.float64)
        nm = C.c_int64(0)
        self._chk(self._lib.glio_assoc_scan_to_map(self._h, C.c_int(slot), p, C.c_int64(n), C.c_int(stride), C.c_int(mem),
                                                   _ptr(t), _ptr(q), C.byref(nm)))
        return int(nm.value)

    def window_set_scans(self, scans):
        args = [_points_arg(s) for s in scans]
        self._scan_keep = [a[0] for a in args]
        W = len(args)
        ptrs = (C.c_void_p * W)(*[a[1] for a in args])
        Q = (C.c_int64 * W)(*[a[2] for a in args])
        strides = {a[3] for a in args}; mems = {a[4] for a in args}
        assert len(strides) == 1 and len(mems) == 1
        self._chk(self._lib.glio_window_set_scans(self._h, C.c_int(W), ptrs, Q, C.c_int(strides.pop()), C.c_int(mems.pop())))

    def window_set_scan(self, slot, scan):
        keep, p, n, stride, mem = _points_arg(scan)
        if not hasattr(self, "_slot_keep"):
            self._slot_keep = {}
        self._slot_keep[slot] = keep
        self._chk(self._lib.glio_window_set_scan(self._h, C.c_int(slot), p, C.c_int64(n), C.c_int(stride), C.c_int(mem)))

    def window_slide(self, W):
        self._chk(self._lib.glio_window_slide(self._h, C.c_int(W)))

    def window_associate(self, poses_body):
        pb = np.ascontiguousarray(poses_body, np.float64).reshape(-1, 7)
        W = len(pb)
        nm = np.zeros(W, np.int64)
        self._chk(self._lib.glio_window_associate(self._h, C.c_int(W), _ptr(pb), _ptr(nm)))
        return nm

    def get_matches(self, slot, capacity):
        cp = np.empty((capacity, 3), np.float32); nsd = np.empty((capacity, 4), np.float32)
        w = np.empty(capacity, np.float32); src = np.empty(capacity, np.int32)
        n = C.c_int64(0)
        self._chk(self._lib.glio_get_matches(self._h, C.c_int(slot), C.c_int64(capacity), _ptr(cp), _ptr(nsd), _ptr(w), _ptr(src), C.byref(n)))
        n = int(n.value)
        return dict(cp=cp[:n], nsd=nsd[:n], weight=w[:n], src=src[:n], n=n)

    def get_assoc_debug(self, slot, Q):
        out = dict(status=np.empty(Q, np.uint8), idx5=np.empty((Q, 5), np.int32), sqd5=np.empty((Q, 5), np.float32),
                   pm=np.empty((Q, 3), np.float32), plane=np.empty((Q, 4), np.float64))
        self._chk(self._lib.glio_get_assoc_debug(self._h, C.c_int(slot), C.c_int64(Q), _ptr(out["status"]), _ptr(out["idx5"]),
                                                 _ptr(out["sqd5"]), _ptr(out["pm"]), _ptr(out["plane"])))
        return out

    def select(self, slot, keep):
        if keep is None:
            self._chk(self._lib.glio_select(self._h, C.c_int(slot), None, C.c_int64(-1)))
            return
        k = np.ascontiguousarray(keep, np.int32)
        self._chk(self._lib.glio_select(self._h, C.c_int(slot), _ptr(k), C.c_int64(len(k))))

    # ---- K2
    def eval_unary(self, poses_body, jac_kind=0, want_jac=True):
        pb = np.ascontiguousarray(poses_body, np.float64).reshape(-1, 7)
        W = len(pb)
        H = np.zeros((W, 6, 6)) if want_jac else None
        g = np.zeros((W, 6)) if want_jac else None
        cost = np.zeros(W)
        self._chk(self._lib.glio_eval_unary(self._h, C.c_int(W), _ptr(pb), C.c_int(jac_kind), _ptr(H), _ptr(g), _ptr(cost)))
        return dict(H=H, g=g, cost=cost)

    def eval_unary_residuals(self, slot, pose_body, capacity, jac_kind=0):
        pb = np.ascontiguousarray(pose_body, np.float64)
        r = np.empty(capacity); J = np.empty((capacity, 6)); n = C.c_int64(0)
        self._chk(self._lib.glio_eval_unary_residuals(self._h, C.c_int(slot), _ptr(pb), C.c_int(jac_kind), C.c_int64(capacity),
                                                      _ptr(r), _ptr(J), C.byref(n)))
        n = int(n.value)
        return r[:n], J[:n]


# ---------------------------------------------------------------------------------------------------
# minimizer iteration (ceres::Solve replacement for the window problem) and stand-in host factors
# ---------------------------------------------------------------------------------------------------
class SolverOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("dogleg_type", C.c_int32), ("use_nonmonotonic_steps", C.c_int32),
                ("max_consecutive_nonmonotonic_steps", C.c_int32), ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("fuse_candidate_jacobian", C.c_int32), ("trust_region_strategy", C.c_int32)]


class Iteration(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32), ("reserved", C.c_int32),
                ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double),
                ("step_norm", C.c_double), ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double), ("mu", C.c_double)]


class SolverSummary(C.Structure):
    _fields_ = [("termination", C.c_int32), ("num_iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("num_evaluations", C.c_int32), ("num_jacobian_evaluations", C.c_int32),
                ("num_linear_solves", C.c_int32), ("num_valid_steps", C.c_int32), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("eval_seconds", C.c_double), ("linear_solver_seconds", C.c_double),
                ("total_seconds", C.c_double), ("message", C.c_char * 128)]


HOST_FACTORS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                              C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


def default_solver_options(**kw):
    o = SolverOptions()
    lib().glio_default_solver_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def iterations_to_dicts(log, n):
    names = [f[0] for f in Iteration._fields_ if f[0] != "reserved"]
    return [{k: getattr(log[i], k) for k in names} for i in range(n)]


class HostFactorSet:
    """Stand-in host factors (prior / between / range) evaluated analytically on the CPU by the library."""

    def __init__(self):
        L = lib()
        L.glio_hf_create.restype = C.c_void_p
        L.glio_hf_destroy.argtypes = [C.c_void_p]
        self._lib = L
        self._h = C.c_void_p(L.glio_hf_create())

    def add_prior(self, kf, t0, q0, sb0, sqrt_w):
        a = [np.ascontiguousarray(v, np.float64) if v is not None else None for v in (t0, q0, sb0, sqrt_w)]
        self._lib.glio_hf_add_prior(self._h, C.c_int(kf), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]))

    def add_between(self, i, j, dp, dq, dv, dt, sqrt_w):
        a = [np.ascontiguousarray(v, np.float64) for v in (dp, dq, dv, sqrt_w)]
        self._lib.glio_hf_add_between(self._h, C.c_int(i), C.c_int(j), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), C.c_double(dt), _ptr(a[3]))

    def add_range(self, kf, lever, sat, rho, w):
        a = [np.ascontiguousarray(v, np.float64) for v in (lever, sat)]
        self._lib.glio_hf_add_range(self._h, C.c_int(kf), _ptr(a[0]), _ptr(a[1]), C.c_double(rho), C.c_double(w))

    def set_marg_prior(self, prior):
        """Attach a MargPrior (or None) to this factor set: it then acts as the MarginalizationFactor of the window."""
        self._prior_keep = prior
        rc = self._lib.glio_hf_set_marg_prior(self._h, prior._h if prior is not None else None)
        assert rc == 0

    def marg_half_bandwidth(self):
        return int(self._lib.glio_hf_marg_half_bandwidth(self._h))

    def marg_evaluate(self, poses, speed_bias):
        """The non-LiDAR part of MarginalizationInfo's A, b (marginalisation ordering, N = 6W + 18)."""
        pb = np.ascontiguousarray(poses, np.float64).reshape(-1, 7); W = len(pb)
        sb = np.ascontiguousarray(speed_bias, np.float64).reshape(W, 9)
        N = 6 * W + 18
        A = np.zeros((N, N)); b = np.zeros(N)
        rc = self._lib.glio_hf_marg_evaluate(self._h, C.c_int(W), _ptr(pb), _ptr(sb), _ptr(A), _ptr(b))
        assert rc == 0
        return A, b

    def evaluate(self, poses, speed_bias=None, want_jac=True):
        pb = np.ascontiguousarray(poses, np.float64).reshape(-1, 7); W = len(pb)
        sb = None if speed_bias is None else np.ascontiguousarray(speed_bias, np.float64).reshape(W, 9)
        n = W * (15 if sb is not None else 6)
        H = np.zeros((n, n)); g = np.zeros(n); c = np.zeros(1)
        rc = self._lib.glio_hf_evaluate(self._h, C.c_int(W), _ptr(pb), _ptr(sb), C.c_int(1 if want_jac else 0), _ptr(H), _ptr(g), _ptr(c))
        assert rc == 0
        return H, g, float(c[0])

    @property
    def callback(self):
        return C.cast(self._lib.glio_hf_evaluate, HOST_FACTORS_FN), self._h

    def __del__(self):
        try:
            if self._h:
                self._lib.glio_hf_destroy(self._h); self._h = None
        except Exception:
            pass


def _extract_features(self, cloud_xyzi, scan_start, scan_end, ds_rate=1, edge_thres=1.0, surf_thres=0.1, ds_v=0.4, intensity_offset=None):
    """Preprocessing::cloudHandler's feature extraction (GLIO/src/Preprocessing.cpp:529-655) on the device."""
    keep, p, n, stride, mem = _points_arg(cloud_xyzi)
    ioff = intensity_offset if intensity_offset is not None else (4 if stride == 8 else 3)
    ss = np.ascontiguousarray(scan_start, np.int32); se = np.ascontiguousarray(scan_end, np.int32); S = len(ss)
    out = dict(curvature=np.empty(n, np.float32), label=np.empty(n, np.int8), sharp=np.empty(12 * S, np.int32), less_sharp=np.empty(60 * S, np.int32),
               flat=np.empty(24 * S, np.int32), less_flat=np.empty(n, np.int32), less_flat_ds=np.empty((n, 4), np.float32))
    cnt = [C.c_int64(0) for _ in range(5)]
    self._chk(self._lib.glio_extract_features(self._h, p, C.c_int64(n), C.c_int(stride), C.c_int(ioff), C.c_int(mem), C.c_int(S), _ptr(ss), _ptr(se), C.c_int(ds_rate),
                                              C.c_double(edge_thres), C.c_double(surf_thres), C.c_float(ds_v), _ptr(out["curvature"]), _ptr(out["label"]),
                                              _ptr(out["sharp"]), C.byref(cnt[0]), _ptr(out["less_sharp"]), C.byref(cnt[1]), _ptr(out["flat"]), C.byref(cnt[2]),
                                              _ptr(out["less_flat"]), C.byref(cnt[3]), _ptr(out["less_flat_ds"]), C.byref(cnt[4])))
    for k, c in zip(("sharp", "less_sharp", "flat", "less_flat", "less_flat_ds"), cnt):
        out[k] = out[k][:c.value]
    return out


Context.extract_features = _extract_features


HOST_MARG_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


class MargPrior:
    """glio_marg_prior: linearized_jacobians / residuals + keep_block_data of MarginalizationInfo (next window's numbering)."""

    def __init__(self, handle, lib_):
        self._h = handle; self._lib = lib_
        n = C.c_int(0); W = C.c_int(0)
        self._lib.glio_marg_prior_size(self._h, C.byref(n), C.byref(W))
        self.n, self.W = n.value, W.value

    @classmethod
    def from_arrays(cls, W, lin_jac, lin_res, x0_pose, x0_sb):
        L = lib()
        a = [np.ascontiguousarray(v, np.float64) for v in (lin_jac, lin_res, x0_pose, x0_sb)]
        h = L.glio_marg_prior_create(C.c_int(W), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]))
        assert h
        return cls(C.c_void_p(h), L)

    def arrays(self):
        n, W = self.n, self.W
        out = dict(W=W, lin_jac=np.zeros((n, n)), lin_res=np.zeros(n), x0_pose=np.zeros((W - 1, 7)), x0_sb=np.zeros(9), A_info=np.zeros((n, n)), b_info=np.zeros(n))
        self._lib.glio_marg_prior_get(self._h, _ptr(out["lin_jac"]), _ptr(out["lin_res"]), _ptr(out["x0_pose"]), _ptr(out["x0_sb"]), _ptr(out["A_info"]), _ptr(out["b_info"]))
        return out

    def __del__(self):
        try:
            if self._h:
                self._lib.glio_marg_prior_destroy(self._h); self._h = None
        except Exception:
            pass


def _window_marginalize(self, poses, speed_bias, host_factors=None, eps=1e-8):
    """MarginalizationInfo::PreMarginalize + Marginalize of KF0 (Estimator.cpp:2462-2608): LiDAR blocks from the device, the
    other factors from `host_factors` (a HostFactorSet).  Returns a MargPrior for the next window."""
    pb = np.ascontiguousarray(poses, np.float64).reshape(-1, 7); W = len(pb)
    sb = np.ascontiguousarray(speed_bias, np.float64).reshape(W, 9)
    if host_factors is None:
        fn, user = C.cast(None, HOST_MARG_FN), None
    else:
        fn, user = C.cast(self._lib.glio_hf_marg_evaluate, HOST_MARG_FN), host_factors._h
    out = C.c_void_p()
    self._chk(self._lib.glio_window_marginalize(self._h, C.c_int(W), _ptr(pb), _ptr(sb), fn, user, C.c_double(eps), C.byref(out)))
    return MargPrior(out, self._lib)


Context.window_marginalize = _window_marginalize


class MargJob:
    """glio_marg_job: a marginalisation whose host half runs on the context's worker thread; wait() joins it -> MargPrior."""

    def __init__(self, ctx, handle, keep):
        self._ctx = ctx; self._h = handle; self._keep = keep          # keep: the arrays / factor set the job may still read

    def wait(self):
        assert self._h is not None, "job already waited for"
        out = C.c_void_p()
        h, self._h = self._h, None
        self._ctx._chk(self._ctx._lib.glio_marg_job_wait(h, C.byref(out)))
        self._keep = None
        return MargPrior(out, self._ctx._lib)


def _window_marginalize_async(self, poses, speed_bias, host_factors=None, eps=1e-8):
    """glio_window_marginalize_async: queues the device half, runs the host callback, returns a MargJob at once."""
    pb = np.ascontiguousarray(poses, np.float64).reshape(-1, 7); W = len(pb)
    sb = np.ascontiguousarray(speed_bias, np.float64).reshape(W, 9)
    if host_factors is None:
        fn, user = C.cast(None, HOST_MARG_FN), None
    else:
        fn, user = C.cast(self._lib.glio_hf_marg_evaluate, HOST_MARG_FN), host_factors._h
    job = C.c_void_p()
    self._chk(self._lib.glio_window_marginalize_async(self._h, C.c_int(W), _ptr(pb), _ptr(sb), fn, user, C.c_double(eps), C.byref(job)))
    return MargJob(self, job, (pb, sb, host_factors))


Context.window_marginalize_async = _window_marginalize_async


def _window_solve(self, poses, speed_bias=None, host_factors=None, options=None, max_log=64, band=None):
    """ceres::Solve for the window problem (Estimator.cpp:2424-2433).  Returns dict(poses, speed_bias, summary, iterations, steps)."""
    pb = np.array(poses, np.float64).reshape(-1, 7).copy(); W = len(pb)
    sb = None if speed_bias is None else np.array(speed_bias, np.float64).reshape(W, 9).copy()
    n = W * (15 if sb is not None else 6)
    opt = options if options is not None else default_solver_options()
    summ = SolverSummary(); log = (Iteration * max_log)(); steps = np.zeros((max_log, n))
    if host_factors is None:
        fn, user = C.cast(None, HOST_FACTORS_FN), None
    elif isinstance(host_factors, HostFactorSet):
        fn, user = host_factors.callback
    else:
        fn, user = host_factors, None          # a HOST_FACTORS_FN instance
    if band is not None and isinstance(host_factors, HostFactorSet):
        # host factors straight into band storage (band = half bandwidth, e.g. 29 for prior + chain with speed-bias)
        bfn = C.cast(self._lib.glio_hf_evaluate_band, HOST_FACTORS_BAND_FN)
        self._chk(self._lib.glio_window_solve_band(self._h, C.c_int(W), _ptr(pb), _ptr(sb), bfn, C.c_int(band), host_factors._h, C.byref(opt),
                                                   C.byref(summ), log, C.c_int(max_log), _ptr(steps), C.c_int64(steps.size)))
    else:
        self._chk(self._lib.glio_window_solve(self._h, C.c_int(W), _ptr(pb), _ptr(sb), fn, user, C.byref(opt), C.byref(summ), log,
                                              C.c_int(max_log), _ptr(steps), C.c_int64(steps.size)))
    return dict(poses=pb, speed_bias=sb, summary=summ, iterations=iterations_to_dicts(log, min(summ.num_iterations, max_log)),
                steps=steps[:summ.num_valid_steps])


Context.window_solve = _window_solve


# ---------------------------------------------------------------------------------------------------
# batch (scan-to-multiscan) path
# ---------------------------------------------------------------------------------------------------
def _batch_set_frame(self, frame, scan_xyz, pose):
    keep, p, n, stride, mem = _points_arg(scan_xyz)
    if not hasattr(self, "_frame_keep"):
        self._frame_keep = {}
    self._frame_keep[frame] = keep
    ps = np.ascontiguousarray(pose, np.float64)
    self._chk(self._lib.glio_batch_set_frame(self._h, C.c_int(frame), p, C.c_int64(n), C.c_int(stride), C.c_int(mem), _ptr(ps)))


def _batch_set_pose(self, frame, pose):
    ps = np.ascontiguousarray(pose, np.float64)
    self._chk(self._lib.glio_batch_set_pose(self._h, C.c_int(frame), _ptr(ps)))


def _batch_associate_pairs(self, cur, oth):
    cur = np.ascontiguousarray(cur, np.int32); oth = np.ascontiguousarray(oth, np.int32)
    nm = np.zeros(len(cur), np.int64)
    self._chk(self._lib.glio_batch_associate_pairs(self._h, _ptr(cur), _ptr(oth), C.c_int64(len(cur)), _ptr(nm)))
    return nm


def _batch_get_matches(self, cur, oth, capacity):
    cp = np.empty((capacity, 3), np.float32); w = np.empty(capacity, np.float32); nc = np.empty((capacity, 6), np.float64)
    src = np.empty(capacity, np.int32); n = C.c_int64(0)
    self._chk(self._lib.glio_batch_get_matches(self._h, C.c_int(cur), C.c_int(oth), C.c_int64(capacity), _ptr(cp), _ptr(w), _ptr(nc), _ptr(src), C.byref(n)))
    n = int(n.value)
    return dict(cp=cp[:n], weight=w[:n], normal_cent=nc[:n], src=src[:n], n=n)


def _batch_select(self, cur, oth, keep):
    if keep is None:
        self._chk(self._lib.glio_batch_select(self._h, C.c_int(cur), C.c_int(oth), None, C.c_int64(-1)))
        return
    k = np.ascontiguousarray(keep, np.int32)
    self._chk(self._lib.glio_batch_select(self._h, C.c_int(cur), C.c_int(oth), _ptr(k), C.c_int64(len(k))))


def _batch_pair_list(self):
    n = C.c_int64(0)
    self._chk(self._lib.glio_batch_pair_list(self._h, C.c_int64(0), None, None, C.byref(n)))
    cur = np.zeros(n.value, np.int32); oth = np.zeros(n.value, np.int32)
    self._chk(self._lib.glio_batch_pair_list(self._h, C.c_int64(n.value), _ptr(cur), _ptr(oth), C.byref(n)))
    return cur, oth


def _batch_clear(self):
    self._chk(self._lib.glio_batch_clear(self._h))


def _eval_binary(self, poses, want_jac=True):
    pb = np.ascontiguousarray(poses, np.float64).reshape(-1, 7); K = len(pb)
    n = C.c_int64(0)
    self._chk(self._lib.glio_batch_pair_list(self._h, C.c_int64(0), None, None, C.byref(n)))
    P = int(n.value)
    Hd = np.zeros((K, 6, 6)) if want_jac else None; Ho = np.zeros((P, 6, 6)) if want_jac else None
    g = np.zeros((K, 6)) if want_jac else None; cost = np.zeros(1)
    self._chk(self._lib.glio_eval_binary(self._h, C.c_int(K), _ptr(pb), _ptr(Hd), _ptr(Ho), _ptr(g), _ptr(cost)))
    return dict(Hdiag=Hd, Hoff=Ho, g=g, cost=float(cost[0]))


HOST_FACTORS_BAND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                   C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


def batch_solver_options(**kw):
    """Estimator.cpp:3275-3281: SUBSPACE_DOGLEG, nonmonotonic steps, max_num_iter (100 in the shipped yaml)."""
    base = dict(dogleg_type=1, use_nonmonotonic_steps=1, max_num_iterations=100)
    base.update(kw)
    return default_solver_options(**base)


def _batch_declare_pairs(self, cur, oth):
    cur = np.ascontiguousarray(cur, np.int32); oth = np.ascontiguousarray(oth, np.int32)
    self._chk(self._lib.glio_batch_declare_pairs(self._h, _ptr(cur), _ptr(oth), C.c_int64(len(cur))))


def _batch_solve(self, poses, speed_bias=None, host_factors=None, options=None, max_log=256):
    pb = np.array(poses, np.float64).reshape(-1, 7).copy(); K = len(pb)
    sb = None if speed_bias is None else np.array(speed_bias, np.float64).reshape(K, 9).copy()
    n = K * (15 if sb is not None else 6)
    opt = options if options is not None else batch_solver_options()
    summ = SolverSummary(); log = (Iteration * max_log)(); steps = np.zeros((max_log, n))
    if host_factors is None:
        fn, user = C.cast(None, HOST_FACTORS_BAND_FN), None
    else:
        fn, user = C.cast(self._lib.glio_hf_evaluate_band, HOST_FACTORS_BAND_FN), host_factors._h
    self._chk(self._lib.glio_batch_solve(self._h, C.c_int(K), _ptr(pb), _ptr(sb), fn, user, C.byref(opt), C.byref(summ), log,
                                         C.c_int(max_log), _ptr(steps), C.c_int64(steps.size)))
    return dict(poses=pb, speed_bias=sb, summary=summ, iterations=iterations_to_dicts(log, min(summ.num_iterations, max_log)),
                steps=steps[:summ.num_valid_steps])


def _set_allreduce(self, fn, user=None):
    """fn: an ALLREDUCE_FN instance (or a raw function pointer from libglio_nccl.so); keeps a reference alive."""
    self._allreduce_keep = (fn, user)
    self._chk(self._lib.glio_set_allreduce(self._h, fn, user))


def _set_edges(self, slot, cp, pa, pb, s):
    cp = np.ascontiguousarray(cp, np.float32).reshape(-1, 3); pa = np.ascontiguousarray(pa, np.float32).reshape(-1, 3)
    pb = np.ascontiguousarray(pb, np.float32).reshape(-1, 3); s = np.ascontiguousarray(s, np.float64)
    self._chk(self._lib.glio_set_edges(self._h, C.c_int(slot), _ptr(cp), _ptr(pa), _ptr(pb), _ptr(s), C.c_int64(len(s))))


def _eval_edge(self, poses_body, want_jac=True):
    pb = np.ascontiguousarray(poses_body, np.float64).reshape(-1, 7); W = len(pb)
    H = np.zeros((W, 6, 6)) if want_jac else None; g = np.zeros((W, 6)) if want_jac else None; cost = np.zeros(W)
    self._chk(self._lib.glio_eval_edge(self._h, C.c_int(W), _ptr(pb), _ptr(H), _ptr(g), _ptr(cost)))
    return dict(H=H, g=g, cost=cost)


Context.set_edges = _set_edges
Context.eval_edge = _eval_edge
def _batch_set_pair_matches(self, cur, oth, cp, normal_cent, weight):
    """Upload the matches of pair (cur, oth) instead of associating them (score = batch_score * weight)."""
    cp = np.ascontiguousarray(cp, np.float32).reshape(-1, 3); nc = np.ascontiguousarray(normal_cent, np.float64).reshape(-1, 6)
    w = np.ascontiguousarray(weight, np.float32)
    self._chk(self._lib.glio_batch_set_pair_matches(self._h, C.c_int(cur), C.c_int(oth), _ptr(cp), _ptr(nc), _ptr(w), C.c_int64(len(w))))


Context.batch_set_pair_matches = _batch_set_pair_matches
Context.batch_declare_pairs = _batch_declare_pairs
Context.batch_solve = _batch_solve
Context.set_allreduce = _set_allreduce
Context.batch_set_frame = _batch_set_frame
Context.batch_set_pose = _batch_set_pose
Context.batch_associate_pairs = _batch_associate_pairs
Context.batch_get_matches = _batch_get_matches
Context.batch_select = _batch_select
Context.batch_pair_list = _batch_pair_list
Context.batch_clear = _batch_clear
Context.eval_binary = _eval_binary
