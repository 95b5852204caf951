"""
Writes tests/golden/ld_zarr_tiny: a tiny LD matrix in the Zarr v2 layout magenpy's LDMatrix keeps on disk (group with
matrix/data = upper-triangular int8 codes without the diagonal, matrix/indptr = int64 row pointers, metadata/*, group
attributes), zlib-compressed chunks, plus ld_zarr_tiny_expected.npz with the arrays it must read back as.
magenpy / zarr are not installable in this image: the store is written with viprs_b200.ingest.write_zarr_1d.
Run from the repository root:  python tests/golden/make_ld_zarr_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from tests_util import make_block_ld          # noqa: E402
from viprs_b200 import ingest                 # noqa: E402

blocks = (40, 25, 60)
P = make_block_ld(np.random.default_rng(20261017), blocks, np.int8, np.float32)
root = os.path.join(HERE, "ld_zarr_tiny")
os.makedirs(root, exist_ok=True)
with open(os.path.join(root, ".zgroup"), "w") as f:
    json.dump({"zarr_format": 2}, f)
with open(os.path.join(root, ".zattrs"), "w") as f:
    json.dump({"Chromosome": 22, "LD estimator": "block", "Sample size": 50000, "Genome build": "GRCh37",
               "Estimator properties": {"LD blocks": [[0, 40], [40, 65], [65, 125]]}}, f)
for sub in ("matrix", "metadata"):
    os.makedirs(os.path.join(root, sub), exist_ok=True)
    with open(os.path.join(root, sub, ".zgroup"), "w") as f:
        json.dump({"zarr_format": 2}, f)
ingest.write_zarr_1d(os.path.join(root, "matrix", "data"), P["data"], 1024, {"id": "zlib", "level": 5})
ingest.write_zarr_1d(os.path.join(root, "matrix", "indptr"), P["indptr"], 64, {"id": "zlib", "level": 5})
ingest.write_zarr_1d(os.path.join(root, "metadata", "bp"), np.arange(P["M"], dtype=np.int32) * 1000 + 16050000, 64, None)
np.savez_compressed(os.path.join(HERE, "ld_zarr_tiny_expected.npz"), data=P["data"], indptr=P["indptr"], beta=P["beta"],
                    blocks=np.array(blocks))
print("wrote", root, P["M"], "SNPs,", P["data"].shape[0], "stored entries")
