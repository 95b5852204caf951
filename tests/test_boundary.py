"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, refuses to compute without a device (no CPU fallback), and the product never touches oracle/."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _built():
    from viprs_b200 import build as vbuild
    return vbuild.build()


def test_library_exports_every_declared_symbol():
    _built()
    import viprs_b200
    from viprs_b200 import _lib
    L = viprs_b200.lib()
    header = open(os.path.join(ROOT, "include", "viprs_b200.h")).read()
    declared = set(re.findall(r"\b(viprs_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 10
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/viprs_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in viprs_b200/_lib.py"
    assert b"sm_100a" in L.viprs_b200_version()
    assert L.viprs_b200_strerror(-5).decode().startswith("no CUDA device")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _built()
    import viprs_b200
    M = 8
    lb = np.arange(1, M + 1, dtype=np.int32)
    ip = np.concatenate([[0], np.cumsum(np.arange(M - 1, -1, -1))]).astype(np.int64)
    data = np.zeros(int(ip[-1]), np.float32)
    z = [np.zeros(M, np.float32) for _ in range(9)]
    with pytest.raises(viprs_b200.ViprsB200Error) as ei:
        viprs_b200.cpp_e_step(lb, ip, data, *z, 1.0, 1, True)
    assert ei.value.code == -5        # VIPRS_B200_ENODEVICE


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]_ref|libviprs_(ref|port)", re.M)
    for d, _, files in os.walk(os.path.join(ROOT, "viprs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert not pat.search(src), f"{os.path.join(d, f)} references the oracle"
    src = open(os.path.join(ROOT, "include", "viprs_b200.h")).read()
    assert not pat.search(src)
