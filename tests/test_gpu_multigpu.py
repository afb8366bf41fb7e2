"""the same mixed-resolution batch on 1 rank and on N ranks (torchrun, NCCL, LPT sharder + ragged gather)
must give bit-identical gathered [V, 35203] matrices and scores.  Skipped on a single-GPU box (run with gpurun --gpus 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n, out, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py"), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_n_rank_gather_equals_one_rank(tmp_path, n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    one, many = str(tmp_path / "one.npz"), str(tmp_path / f"r{n}.npz")
    _run(1, one, 29611)
    _run(n, many, 29611 + n)
    a, b = np.load(one), np.load(many)
    assert b["plan"].tolist() != a["plan"].tolist() and len(b["plan"]) == n and len(set(b["plan"].tolist())) > 1     # a ragged split
    assert a["feats"].shape == (7, 35203) and np.array_equal(a["feats"], b["feats"], equal_nan=True)
    assert np.array_equal(a["score"], b["score"])
