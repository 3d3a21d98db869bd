"""K3 (host): the Schur / eigen-decomposition step of MarginalizationInfo::Marginalize vs numpy (CPU only)."""
import ctypes

import numpy as np


def test_marginalize_matches_numpy_invariants():
    from glio_b200 import api
    lib = api.lib()
    rng = np.random.default_rng(0)
    N, m = 138, 15                       # 6W+18 at W = 20, drop KF0's (t, q, speed-bias)
    J = rng.normal(size=(400, N)); J[:, 20:40] *= 1e-3
    J[:, 100:103] = 0.0                  # unobservable directions -> eigenvalues below eps are truncated
    A = J.T @ J; b = J.T @ rng.normal(size=400)
    n = N - m
    LJ = np.zeros((n, n)); lr = np.zeros(n)
    rc = lib.glio_marginalize(A.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(N), ctypes.c_int(m),
                              ctypes.c_double(1e-8), LJ.ctypes.data_as(ctypes.c_void_p), lr.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    # numpy restatement of MarginalizationFactor.cpp:176-201
    eps = 1e-8
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    Ainv = V @ np.diag(np.where(w > eps, 1 / np.where(w > eps, w, 1), 0)) @ V.T
    Ar = A[m:, m:] - A[m:, :m] @ Ainv @ A[:m, m:]
    br = b[m:] - A[m:, :m] @ Ainv @ b[:m]
    w2, V2 = np.linalg.eigh(Ar)
    S = np.where(w2 > eps, w2, 0); Si = np.where(w2 > eps, 1 / np.where(w2 > eps, w2, 1), 0)
    Jref = np.diag(np.sqrt(S)) @ V2.T; rref = np.diag(np.sqrt(Si)) @ V2.T @ br
    assert (w2 <= eps).sum() >= 3        # the truncation branch is exercised
    scale = np.abs(Ar).max()
    assert np.max(np.abs(LJ.T @ LJ - Jref.T @ Jref)) <= 1e-9 * scale          # J^T J (invariant to eigenvector signs/order)
    assert np.max(np.abs(LJ.T @ lr - Jref.T @ rref)) <= 1e-9 * np.abs(Jref.T @ rref).max()
    assert abs(lr @ lr - rref @ rref) <= 1e-9 * (rref @ rref)
