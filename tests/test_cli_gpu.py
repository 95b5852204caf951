"""
bin/viprs_b200_fit: the thin driver over the path -- LD store read by viprs_b200.ingest,
summary statistics table, EM and grid-search fits, the reference's output tables -- on the committed tiny LD store.
The fit table must hold exactly what the model classes produce through the Python API on the same inputs.
"""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
FIXTURE = os.path.join(GOLDEN, "ld_zarr_tiny")
CLI = os.path.join(ROOT, "bin", "viprs_b200_fit")


def _sumstats(path):
    ref = np.load(os.path.join(GOLDEN, "ld_zarr_tiny_expected.npz"))
    M = int(ref["blocks"].sum())
    n = 50000.0
    r = ref["beta"].astype(np.float64)                    # standardized marginal betas of the fixture
    z = r * np.sqrt((n - 2.0) / (1.0 - r * r))            # inverse of z / sqrt(n - 2 + z^2)
    with open(path, "w") as f:
        f.write("CHR\tSNP\tPOS\tA1\tA2\tN\tZ\n")
        for j in range(M):
            f.write(f"22\trs{j}\t{16050000 + 1000 * j}\tA\tG\t{int(n)}\t{z[j]:.12g}\n")
    return M, r, n


def _read_table(path):
    import pandas as pd
    with gzip.open(path, "rt") as f:
        return pd.read_csv(f, sep="\t")


def test_cli_em_matches_api(tmp_path):
    import pandas as pd
    import torch
    from viprs_b200 import ingest
    from viprs_b200.model import VIPRS
    ss = str(tmp_path / "ss.tsv")
    M, r, n = _sumstats(ss)
    out = tmp_path / "out"
    subprocess.run([sys.executable, CLI, "-l", FIXTURE, "-s", ss, "--output-dir", str(out), "--max-iter", "30", "--genomewide"],
                   check=True, timeout=600)
    fit = _read_table(str(out / "VIPRS_EM.fit.gz"))
    hyp = pd.read_csv(str(out / "VIPRS_EM.hyp"), sep="\t")
    assert list(fit.columns) == ["CHR", "SNP", "POS", "A1", "A2", "BETA", "PIP", "VAR_BETA"] and len(fit) == M
    assert {"ELBO", "Residual_variance", "Heritability", "Proportion_causal"} <= set(hyp["Parameter"])
    np.random.seed(7209)
    data = ingest.data_from_zarr({22: FIXTURE}, {22: r.astype(np.float32)}, {22: np.full(M, n)})
    m = VIPRS(data=data, float_precision="float32")
    m.fit(max_iter=30, device_loop=True, check_every=8)
    assert np.allclose(fit["BETA"].to_numpy(), m.post_mean_beta[22], rtol=1e-4, atol=1e-9)
    assert np.allclose(fit["PIP"].to_numpy(), m.pip[22], rtol=1e-4, atol=1e-9)
    torch.cuda.synchronize()


def test_cli_grid_search_writes_validation(tmp_path):
    import pandas as pd
    ss = str(tmp_path / "ss.tsv")
    M, _, _ = _sumstats(ss)
    out = tmp_path / "out"
    subprocess.run([sys.executable, CLI, "-l", FIXTURE, "-s", ss, "--output-dir", str(out), "--max-iter", "100",
                    "--hyp-search", "GS", "--pi-steps", "3", "--sigma-epsilon-steps", "2", "--genomewide"], check=True, timeout=600)
    fit = _read_table(str(out / "VIPRS_GS.fit.gz"))
    assert len(fit) == M and np.all(np.isfinite(fit["BETA"])) and np.all((fit["PIP"] >= 0) & (fit["PIP"] <= 1))
    val = pd.read_csv(str(out / "VIPRS_GS.validation"), sep="\t")
    assert len(val) == 6 and "ELBO" in val.columns
