"""N>1 host logic on CPU (gloo, world_size 2): keyframe-sharded binary blocks summed over ranks equal the unsharded
evaluation, with the block layout / ownership rules of glio_b200/dist.py."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist_
    sys.path.insert(0, ROOT)
    from glio_b200 import dist, synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist_.init_process_group("gloo", rank=rank, world_size=world)
    K, sr = 8, 2
    B = synth.batch_problem(K=K, Q=1200, seed=40, search_range=sr, rng_range=4.0)
    cur, oth = dist.batch_pairs(K, sr)
    own = dist.owner_of(cur, K, world) == rank
    need = dist.frames_needed(cur, oth, own)
    kc, ko, cp, nc, sc = [], [], [], [], []
    for c, o in zip(cur[own], oth[own]):
        assert c in need and o in need
        r = po.assoc_pair(B["scans"][c], B["poses_init"][c, :3], B["poses_init"][c, 3:], B["scans"][o], B["poses_init"][o, :3], B["poses_init"][o, 3:])
        v = r["status"] == 0
        kc.append(np.full(v.sum(), c, np.int32)); ko.append(np.full(v.sum(), o, np.int32)); cp.append(B["scans"][c][v]); nc.append(r["normal_cent"][v]); sc.append(r["score"][v])
    e = po.eval_binary(B["poses_init"], np.concatenate(kc), np.concatenate(ko), np.concatenate(cp), np.concatenate(nc), np.concatenate(sc), per_residual=False)
    H = torch.from_numpy(e["H"].copy()); g = torch.from_numpy(e["g"].copy()); c = torch.tensor([e["cost_total"]])
    dist_.all_reduce(H); dist_.all_reduce(g); dist_.all_reduce(c)
    if rank == 0:
        np.savez(out, H=H.numpy(), g=g.numpy(), c=c.numpy(), n_own=int(own.sum()), n_need=len(need))
    dist_.destroy_process_group()


def test_sharded_binary_blocks_sum_to_full(tmp_path, oracle):
    sys.path.insert(0, ROOT)
    from glio_b200 import dist, synth
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, 29000 + os.getpid() % 2000, out), nprocs=2, join=True)
    r = np.load(out)
    K, sr = 8, 2
    B = synth.batch_problem(K=K, Q=1200, seed=40, search_range=sr, rng_range=4.0)
    cur, oth = dist.batch_pairs(K, sr)
    kc, ko, cp, nc, sc = [], [], [], [], []
    for c, o in zip(cur, oth):
        a = oracle.assoc_pair(B["scans"][c], B["poses_init"][c, :3], B["poses_init"][c, 3:], B["scans"][o], B["poses_init"][o, :3], B["poses_init"][o, 3:])
        v = a["status"] == 0
        kc.append(np.full(v.sum(), c, np.int32)); ko.append(np.full(v.sum(), o, np.int32)); cp.append(B["scans"][c][v]); nc.append(a["normal_cent"][v]); sc.append(a["score"][v])
    e = oracle.eval_binary(B["poses_init"], np.concatenate(kc), np.concatenate(ko), np.concatenate(cp), np.concatenate(nc), np.concatenate(sc), per_residual=False)
    assert np.max(np.abs(r["H"] - e["H"])) <= 1e-12 * np.max(np.abs(e["H"]))
    assert np.max(np.abs(r["g"] - e["g"])) <= 1e-12 * np.max(np.abs(e["g"]))
    assert r["c"][0] == pytest.approx(e["cost_total"], rel=1e-13)
    assert r["n_own"] == len(cur) // 2 and r["n_need"] < K         # a rank holds its range plus the halo, not everything
