import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_golden(name):
    """-> (dict of arrays, {chrom: {ld_data, ld_indptr, ld_left_bound, std_beta, n_per_snp}})"""
    z = np.load(os.path.join(GOLDEN, name))
    d = {k: z[k] for k in z.files}
    chroms = {}
    for k, v in d.items():
        if k.startswith("in_"):
            _, c, field = k.split("_", 2)
            chroms.setdefault(int(c), {})[field] = v
    return d, chroms


def relmax(a, b):
    """max|a-b| / max|b| -- the max-norm-relative metric of SURVEY.md section 8c (most eta_j are ~0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


@pytest.fixture(scope="session")
def oracle_built():
    import oracle
    from oracle import cpu
    if not os.path.exists(cpu._path("port")):
        cpu.build()
    return oracle
