"""The C-ABI library loads on a CPU-only box and exports every symbol include/glio_b200.h declares; the product path
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch  # noqa: F401  (first: torch must bind its bundled CUDA runtime before libglio_b200.so pulls in the system libcudart)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "glio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(glio_[a-z0-9_]+)\s*\(", src))
    names -= {n for n in names if re.search(r"\(\s*\*\s*" + n + r"\s*\)", src)}      # function-pointer typedefs
    return sorted(names)


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "glio_b200", "libglio_b200.so"))
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_nccl_helper_library_loads():
    lib = ctypes.CDLL(os.path.join(ROOT, "glio_b200", "libglio_nccl.so"))
    for n in ("glio_nccl_get_unique_id", "glio_nccl_comm_create", "glio_nccl_allreduce", "glio_nccl_comm_destroy"):
        assert hasattr(lib, n)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from glio_b200 import api
    with pytest.raises(api.GlioError, match="no CUDA device"):
        api.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "glio_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                low = txt.lower()
                assert "pyoracle" not in low and "glio_oracle" not in low and "libglio_oracle" not in low and "go_" + "eval" not in low, f"{f} references the oracle"
                assert "oracle" not in low, f"{f} mentions the oracle (comments included: keep the product tree free of it)"


def test_lidar_pose_matches_numpy():
    from glio_b200 import api, synth
    lib = api.lib()
    prm = api.default_params(q_lb=synth.quat_from_rpy(0.01, -0.02, 0.03), t_lb=[0.1, 0.0, 0.28])
    pose = np.array([3.0, -1.0, 0.2, *synth.quat_from_rpy(0.1, 0.2, -0.4)])
    t2 = np.zeros(3); q2 = np.zeros(4)
    lib.glio_lidar_pose(ctypes.byref(prm), pose.ctypes.data_as(ctypes.c_void_p), t2.ctypes.data_as(ctypes.c_void_p), q2.ctypes.data_as(ctypes.c_void_p))
    tr, qr = synth.lidar_pose_in_world(pose[:3], pose[3:], np.array(list(prm.q_lb)), np.array(list(prm.t_lb)))
    assert np.allclose(t2, tr, atol=1e-14) and np.allclose(q2, qr, atol=1e-14)


def test_batch_pair_enumeration_matches_reference_rule():
    from glio_b200 import dist
    cur, oth = dist.batch_pairs(20, 3)
    assert len(cur) == 20 * 6
    for idx in range(20):
        js = sorted(oth[cur == idx].tolist())
        s = idx - 3 if 3 <= idx < 16 else (0 if idx < 3 else 13)        # Estimator.cpp:3009-3017
        assert js == [j for j in range(s, s + 7) if j != idx]
    own = dist.owner_of(cur, 20, 4)
    assert sorted(set(own.tolist())) == [0, 1, 2, 3] and (np.diff(own[np.argsort(cur, kind="stable")]) >= 0).all()


def test_null_context_is_an_argument_error_not_a_crash():
    """Error convention of the ABI: every entry point returns a negative glio_status, nothing aborts."""
    lib = ctypes.CDLL(os.path.join(ROOT, "glio_b200", "libglio_b200.so"))
    n = ctypes.c_void_p(None)
    z = ctypes.c_int64(0)
    calls = [
        ("glio_set_map", (n, n, z, ctypes.c_int(3), ctypes.c_int(0))),
        ("glio_window_set_scans", (n, ctypes.c_int(1), n, n, ctypes.c_int(3), ctypes.c_int(0))),
        ("glio_window_associate", (n, ctypes.c_int(1), n, n)),
        ("glio_eval_unary", (n, ctypes.c_int(1), n, ctypes.c_int(0), n, n, n)),
        ("glio_eval_binary", (n, ctypes.c_int(1), n, n, n, n, n)),
        ("glio_batch_set_pair_matches", (n, ctypes.c_int(0), ctypes.c_int(1), n, n, n, z)),
        ("glio_localmap_clear", (n,)),
        ("glio_localmap_push", (n, ctypes.c_int(0), n, z, ctypes.c_int(3), ctypes.c_int(0), n, n)),
        ("glio_localmap_pop_front", (n,)),
        ("glio_localmap_build", (n, ctypes.c_float(0.4), n)),
        ("glio_get_map", (n, z, n, n)),
        ("glio_synchronize", (n,)),
    ]
    for name, args in calls:
        f = getattr(lib, name); f.restype = ctypes.c_int
        assert f(*args) == -3, name          # GLIO_ERR_ARG


def test_band_callback_rejects_a_too_narrow_band():
    """a host Hessian entry outside the declared half bandwidth must be an error, not silently dropped (a wrong
    band would give a wrong J^T J with no sign of it)."""
    import ctypes as C
    import numpy as np
    from glio_b200 import api, synth
    rng = np.random.default_rng(0)
    T = synth.trajectory(3, rng)
    hf = api.HostFactorSet()
    sw = np.concatenate([np.full(3, 20.0), np.full(3, 50.0), np.full(9, 5.0)])
    hf.add_between(0, 1, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3), 0.1, sw)
    L = api.lib()
    n = 45
    sb = np.zeros((3, 9)); g = np.zeros(n); cost = np.zeros(1)
    for hb, want_ok in ((29, True), (10, False)):
        Hb = np.zeros(n * (hb + 1))
        rc = L.glio_hf_evaluate_band(hf._h, C.c_int(3), T.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p), C.c_int(1),
                                     Hb.ctypes.data_as(C.c_void_p), C.c_int(hb), g.ctypes.data_as(C.c_void_p), cost.ctypes.data_as(C.c_void_p))
        assert (rc == 0) == want_ok, (hb, rc)


def test_bench_splits_a_socket_between_its_ranks_by_whole_cores():
    """bench.py binds every rank to the CPUs local to its GPU; ranks behind the same socket get disjoint physical cores"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    sib = lambda c: (c % 64, c % 64 + 64)                                  # 2 hardware threads per core: i and i + 64
    socket0 = set(range(0, 32)) | set(range(64, 96))
    parts = [b.share_of_cpus(socket0, 4, i, sib) for i in range(4)]
    assert set().union(*parts) == socket0 and sum(len(p) for p in parts) == len(socket0)
    for p in parts:
        assert len(p) == 16 and all(set(sib(c)) <= p for c in p)
    assert b.share_of_cpus({0, 1, 2}, 4, 1, sib) == {0, 1, 2}             # too few CPUs to split: left alone
    assert b.share_of_cpus(socket0, 1, 0, sib) == socket0
    assert b.share_of_cpus(socket0, 3, 2, None) == set(sorted(socket0)[42:])   # no topology information: plain thirds
