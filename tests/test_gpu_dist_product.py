"""The PRODUCT's multi-GPU path (keyframe-sharded batch solve, one all-reduce of the pose-block buffers per
evaluation) under the driver's single-GPU `pytest -m gpu`: two ranks (two processes, torchrun, rendezvous on 127.0.0.1) share
the one GPU, each owns half of the keyframes' pairs, and the library's glio_allreduce_fn hook sums the K x 28 / P x 36 buffers
through torch.distributed.  The sharded run must reproduce the single-rank run: residual count, block values, cost, every
iterate."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "batch_shard_worker.py")


def _run(tmp_path, world, tag):
    out = str(tmp_path / tag)
    env = dict(os.environ); env["OMP_NUM_THREADS"] = "4"
    if world == 1:
        subprocess.check_call([sys.executable, WORKER, out], env=env, timeout=600)
    else:
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                               "--master-port", "29613", WORKER, out], env=env, timeout=600)
    return [json.load(open(out + f".rank{r}.json")) for r in range(world)]


def test_sharded_batch_solve_equals_single_rank(tmp_path):
    one = _run(tmp_path, 1, "one")[0]
    two = _run(tmp_path, 2, "two")
    assert two[0]["hook_calls"] > 0 and two[0]["hook_calls"] == two[1]["hook_calls"]          # the collective really ran, in lockstep
    assert 0 < two[0]["n_match_own"] < one["n_match_own"] and two[0]["n_match_own"] + two[1]["n_match_own"] == one["n_match_own"]
    for r in two:
        assert r["n_match_total"] == one["n_match_total"]
        # blocks after the all-reduce: the sum over ranks equals the single-rank blocks (different summation order only)
        assert np.allclose(r["eval_cost"], one["eval_cost"], rtol=1e-12, atol=0)
        assert np.allclose(np.array(r["eval_g"]), np.array(one["eval_g"]), rtol=1e-10, atol=1e-9)
        assert np.allclose(np.array(r["eval_Hdiag"]), np.array(one["eval_Hdiag"]), rtol=1e-10, atol=1e-9)
        assert r["iterations"] == one["iterations"] >= 3
        assert abs(r["final_cost"] - one["final_cost"]) <= 1e-9 * abs(one["final_cost"])
        for a, b in zip(r["steps"], one["steps"]):
            a = np.array(a).reshape(-1, 6); b = np.array(b).reshape(-1, 6)
            assert np.max(np.abs(a[:, :3] - b[:, :3])) <= 1e-6 and np.max(2 * np.linalg.norm(a[:, 3:] - b[:, 3:], axis=1)) <= 1e-8
        assert np.max(np.abs(np.array(r["poses"]) - np.array(one["poses"]))) <= 1e-8
    assert np.array_equal(np.array(two[0]["poses"]), np.array(two[1]["poses"]))                  # replicated solve: bit-identical on both ranks
