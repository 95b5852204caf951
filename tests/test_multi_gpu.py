"""
Multi-GPU parity: the EM loop on two GPUs -- whole LD blocks sharded across ranks, one NCCL
all-reduce of the M-step / ELBO sums per iteration -- must reproduce the single-GPU run on the same data.  Needs two
CUDA devices (skipped otherwise); the host-side logic of the same path is covered on CPU by the gloo tests in
test_em_host.py.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HELPER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "mgpu_em_history.py")


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two CUDA devices")
def test_two_gpu_em_history_matches_one_gpu(tmp_path):
    one, two = str(tmp_path / "h1.json"), str(tmp_path / "h2.json")
    subprocess.run([sys.executable, HELPER, one], check=True, timeout=600)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                    "127.0.0.1", "--master-port", "29541", HELPER, two], check=True, timeout=600)
    a, b = np.array(json.load(open(one))), np.array(json.load(open(two)))
    assert a.shape == b.shape == (8, 5)
    # float64 sums of per-rank float64 partials: the only difference is the order of a handful of additions
    assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-30)) <= 1e-12
