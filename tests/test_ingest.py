"""
LD ingestion (viprs_b200/ingest.py): the Zarr v2 reader and its codecs on CPU, the committed tiny magenpy-style store
(tests/golden/ld_zarr_tiny, written by tests/golden/make_ld_zarr_fixture.py), and -- on the GPU -- the store read
straight into a DeviceLD and swept against the oracle.
"""
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from tests_util import make_block_ld

FIXTURE = os.path.join(GOLDEN, "ld_zarr_tiny")


def _blosc_frame(raw, typesize, codec, shuffle, blocksize, split=False):
    """A Blosc 1 frame built by hand from the format description (test-side encoder)."""
    import pyarrow as pa
    nbytes = len(raw)
    nblocks = (nbytes + blocksize - 1) // blocksize
    comp_id = {"lz4": 1, "zlib": 3, "zstd": 4}[codec]
    flags = (1 if shuffle else 0) | (0 if split else 0x10) | (comp_id << 5)
    body, bstarts = b"", []
    base = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = raw[b * blocksize:(b + 1) * blocksize]
        if shuffle and typesize > 1:
            ne = len(blk) // typesize
            arr = np.frombuffer(blk, dtype=np.uint8, count=ne * typesize).reshape(ne, typesize).T.reshape(-1)
            blk = arr.tobytes() + blk[ne * typesize:]
        leftover = len(blk) != blocksize
        nsplits = typesize if (split and not leftover and blocksize // typesize >= 128) else 1
        ne = len(blk) // nsplits
        bstarts.append(base + len(body))
        for s in range(nsplits):
            part = blk[s * ne:(s + 1) * ne]
            if codec == "zlib":
                c = zlib.compress(part)
            else:
                c = pa.compress(part, codec={"lz4": "lz4_raw", "zstd": "zstd"}[codec], asbytes=True)
            if len(c) >= len(part):
                c = part                                    # stored raw
            body += struct.pack("<i", len(c)) + c
    hdr = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, base + len(body))
    return hdr + struct.pack("<%di" % nblocks, *bstarts) + body


@pytest.mark.parametrize("codec,shuffle,split,typesize", [("zstd", False, False, 1), ("zstd", True, False, 8), ("lz4", True, True, 8),
                                                          ("zlib", True, False, 2), ("lz4", False, False, 1)])
def test_blosc_frames_decode(codec, shuffle, split, typesize):
    from viprs_b200 import ingest
    rng = np.random.default_rng(1)
    raw = (rng.integers(0, 7, 5000 * typesize).astype(np.uint8) if typesize == 1 else
           np.cumsum(rng.integers(0, 50, 5000)).astype({8: np.int64, 2: np.int16}[typesize]).view(np.uint8)).tobytes()
    frame = _blosc_frame(raw, typesize, codec, shuffle, blocksize=2048 * typesize, split=split)
    assert ingest.blosc_decompress(frame) == raw
    memcpy = struct.pack("<BBBBIII", 2, 1, 0x02, typesize, len(raw), len(raw), 16 + len(raw)) + raw
    assert ingest.blosc_decompress(memcpy) == raw


@pytest.mark.parametrize("compressor", [None, {"id": "zlib", "level": 1}, {"id": "zstd", "level": 3}])
def test_zarr_array_round_trip(tmp_path, compressor):
    from viprs_b200 import ingest
    rng = np.random.default_rng(2)
    for dt, n, chunk in ((np.int8, 10007, 4096), (np.int64, 301, 128), (np.float32, 5, 16)):
        a = rng.integers(-100, 100, n).astype(dt)
        p = str(tmp_path / f"arr_{np.dtype(dt).name}")
        ingest.write_zarr_1d(p, a, chunk, compressor)
        z = ingest.ZarrArray1D(p)
        assert z.n == n and z.dtype == np.dtype(dt)
        assert np.array_equal(z.read(), a)


def test_numcodecs_style_codecs():
    """zstd / lz4 (int32 size prefix + LZ4 block) / bz2 / lzma / gzip chunk payloads as numcodecs writes them."""
    import bz2
    import lzma
    import pyarrow as pa
    from viprs_b200 import ingest
    raw = bytes(range(256)) * 40
    assert ingest.decode_chunk(pa.compress(raw, codec="zstd", asbytes=True), {"id": "zstd"}, len(raw)) == raw
    lz = struct.pack("<i", len(raw)) + pa.compress(raw, codec="lz4_raw", asbytes=True)
    assert ingest.decode_chunk(lz, {"id": "lz4"}, len(raw)) == raw
    assert ingest.decode_chunk(bz2.compress(raw), {"id": "bz2"}, len(raw)) == raw
    assert ingest.decode_chunk(lzma.compress(raw), {"id": "lzma"}, len(raw)) == raw
    assert ingest.decode_chunk(zlib.compress(raw), {"id": "zlib"}, len(raw)) == raw
    import gzip
    assert ingest.decode_chunk(gzip.compress(raw), {"id": "gzip"}, len(raw)) == raw


def test_committed_ld_store_reads_back():
    from viprs_b200 import ingest
    d = ingest.read_ld_zarr(FIXTURE, device="cpu", to_device_ld=False)
    ref = np.load(os.path.join(GOLDEN, "ld_zarr_tiny_expected.npz"))
    assert np.array_equal(d["ld_data"].numpy(), ref["data"]) and np.array_equal(d["ld_indptr"].numpy(), ref["indptr"])
    assert np.array_equal(d["ld_left_bound"].numpy(), np.arange(1, d["n_snps"] + 1))
    assert d["dq_scale"] == 1.0 / 127 and d["attrs"]["LD estimator"] == "block"


def test_bad_stores_are_refused(tmp_path):
    from viprs_b200 import ingest
    P = make_block_ld(np.random.default_rng(0), (10, 5), np.int8, np.float32)
    ingest.write_zarr_1d(str(tmp_path / "s" / "matrix" / "data"), P["data"], 64)
    ip = P["indptr"].copy()
    ip[-1] += 3                                             # does not match matrix/data
    ingest.write_zarr_1d(str(tmp_path / "s" / "matrix" / "indptr"), ip, 64)
    with pytest.raises(ValueError):
        ingest.read_ld_zarr(str(tmp_path / "s"), device="cpu", to_device_ld=False)


@pytest.mark.gpu
def test_store_to_device_and_sweep(oracle_built):
    import torch
    import viprs_b200 as vb
    from viprs_b200 import ingest
    d = ingest.read_ld_zarr(FIXTURE, device="cuda")
    ld = d["ld"]
    ref = np.load(os.path.join(GOLDEN, "ld_zarr_tiny_expected.npz"))
    assert ld.M == d["n_snps"] and ld.n_blocks == len(ref["blocks"]) and ld.max_block == int(ref["blocks"].max())
    M, T = ld.M, np.float32
    rng = np.random.default_rng(4)
    n = np.floor(rng.uniform(4e4, 6e4, M))
    pi, se = 0.05, 0.8
    tau = pi * M / (1 - se)
    vt = n / se + tau
    ul = (np.log(pi) - np.log(1 - pi) + .5 * (np.log(tau) - np.log(vt))).astype(T)
    sv, mm = np.sqrt(.5 * vt).astype(T), (n / (vt * se)).astype(T)
    beta = ref["beta"].astype(T)
    st = {k: np.zeros(M, T) for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.full(M, pi, T)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in st.items()}
    c = lambda a: torch.from_numpy(a).cuda()
    for _ in range(3):
        oracle_built.e_step(np.arange(1, M + 1, dtype=np.int32), ref["indptr"], ref["data"], beta, st["var_gamma"], st["var_mu"],
                            st["eta"], st["q"], st["eta_diff"], ul, sv, mm, d["dq_scale"], 1, True)
        vb.e_step_device(ld, c(beta), dev["var_gamma"], dev["var_mu"], dev["eta"], dev["q"], dev["eta_diff"], c(ul), c(sv), c(mm),
                         d["dq_scale"], True)
    for k in ("eta", "var_gamma", "var_mu", "q"):
        assert relmax(dev[k].cpu().numpy(), st[k]) <= 1e-4, k
