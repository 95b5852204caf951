"""Small numpy helpers shared by the tests (no product or oracle code)."""
import numpy as np


def make_block_ld(rng, blocks, ld_dtype, T, symmetric=False, k=16, alpha=0.5, n=5e4, effect=0.006, p_causal=0.02):
    """Block-diagonal PD LD in magenpy's CSR-without-column-indices layout (upper-triangular, or the
    symmetric `low_memory=False` layout with the unit diagonal)."""
    ld_dtype = np.dtype(ld_dtype)
    data, lens, lbs, betas = [], [], [], []
    row0 = 0
    for B in blocks:
        Z = rng.standard_normal((B, k))
        C = Z @ Z.T / k + 1e-3 * np.eye(B)
        dinv = 1. / np.sqrt(np.diag(C))
        R = alpha * C * dinv[:, None] * dinv[None, :]
        np.fill_diagonal(R, 1.)
        betas.append(R @ (rng.standard_normal(B) * effect * (rng.random(B) < p_causal)) + rng.standard_normal(B) / np.sqrt(n))
        if ld_dtype == np.int8:
            Rq = np.rint(R * 127.).astype(np.int8)
        elif ld_dtype == np.int16:
            Rq = np.rint(R * 32767.).astype(np.int16)
        else:
            Rq = R.astype(ld_dtype)
        if symmetric:
            data.append(Rq.reshape(-1))
            lens.append(np.full(B, B, np.int64))
            lbs.append(np.full(B, row0, np.int32))
        else:
            data.append(Rq[np.triu_indices(B, 1)])
            lens.append(np.arange(B - 1, -1, -1, dtype=np.int64))
            lbs.append(np.arange(row0 + 1, row0 + B + 1, dtype=np.int32))
        row0 += B
    indptr = np.zeros(row0 + 1, np.int64)
    indptr[1:] = np.cumsum(np.concatenate(lens))
    dq = {np.dtype(np.int8): 1. / 127, np.dtype(np.int16): 1. / 32767}.get(ld_dtype, 1.)
    return {"M": row0, "data": np.concatenate(data), "indptr": indptr, "lb": np.concatenate(lbs),
            "beta": np.concatenate(betas).astype(T), "dq": dq, "blocks": list(blocks)}
