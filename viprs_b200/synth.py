"""
Seeded synthetic inputs of the shapes BASELINE.json names: block-diagonal,
positive-definite LD in magenpy's upper-triangular CSR-without-column-indices layout plus
summary statistics.  Pure torch, so the same code generates small cases on the CPU for the parity
tests and the genome-wide cases directly in HBM for bench.py.

Per block b (size B_b):  Z ~ N(0,1)^{B_b x k}, C = Z Z'/k, R = (1-a) I + a D^-1/2 C D^-1/2
(unit diagonal, eigenvalues >= 1-a).  Rows are emitted upper-triangular: row j stores columns
j+1 .. block_end-1, left_bound[j] = j+1, the last row of a block is empty.
Quantised encodings follow VIPRS.py:203-205 (dequantize_scale = 1/iinfo.max):
int8 = rint(127 R), int16 = rint(32767 R).
"""
import math

import numpy as np
import torch

SEED = 7209   # the reference's default seed (bin/viprs_fit:996)

_LD_TORCH = {"float32": torch.float32, "float64": torch.float64, "int8": torch.int8, "int16": torch.int16}


def block_sizes_for(M, block=4096):
    """Equal blocks of `block` SNPs (last one shorter)."""
    sizes = [block] * (M // block)
    if M % block:
        sizes.append(M % block)
    return sizes


def ldetect_like_sizes(M, lo=400, hi=1200, seed=SEED):
    """LDetect-like block sizes ~ U[lo, hi] summing to M (config 1 shape)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    sizes, tot = [], 0
    while tot < M:
        s = int(rng.integers(lo, hi + 1))
        s = min(s, M - tot)
        sizes.append(s)
        tot += s
    return sizes


def dequantize_scale(ld_dtype):
    if ld_dtype == "int8":
        return 1.0 / 127.0
    if ld_dtype == "int16":
        return 1.0 / 32767.0
    return 1.0


def _block_draws(seed, block_id, B, k):
    """The random numbers of LD block `block_id`: numpy Philox keyed on (seed, block_id), so that ANY subset of blocks
    -- a CPU slice for the reference arm, one rank's shard of a multi-GPU run -- reproduces exactly the arrays of the
    full genome, wherever the dense algebra below then runs."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), int(block_id)]))
    return dict(Z=rng.standard_normal((B, k)), u=rng.random(B), b=rng.standard_normal(B), e1=rng.standard_normal(B),
                e2=rng.standard_normal(k))


def make_inputs(block_sizes, ld_dtype="float32", float_dtype=torch.float32, device="cpu", seed=SEED,
                k=64, alpha=0.5, n=300000, h2=0.3, p_causal=0.01, symmetric=False, block_ids=None, M_total=None):
    """
    Returns a dict of torch tensors on `device`:
      ld_data, ld_indptr (int64), ld_left_bound (int32), std_beta, n_per_snp, beta_true
    `symmetric=True` emits the `low_memory=False` layout instead (full block rows incl. diagonal).
    `block_ids` (default 0 .. len-1) name the blocks for the per-block random streams; `M_total` is the genome size
    the effect-size variance refers to (default: the sum of `block_sizes`) -- pass both to generate a shard.
    """
    dev = torch.device(device)
    M = int(sum(block_sizes))
    udt = _LD_TORCH[ld_dtype]
    data, lens, lbs, betas, trues = [], [], [], [], []
    row0 = 0
    sb2 = h2 / (p_causal * (M_total or M))
    ids = list(range(len(block_sizes))) if block_ids is None else list(block_ids)
    f64 = lambda a: torch.from_numpy(a).to(dev)
    for B, bid in zip(block_sizes, ids):
        dr = _block_draws(seed, bid, B, k)
        Z = f64(dr["Z"])
        C = (Z @ Z.T) / k
        dinv = torch.rsqrt(torch.diagonal(C))
        R = alpha * (C * dinv[:, None] * dinv[None, :])
        R.diagonal().add_(1.0 - alpha)
        R.diagonal().fill_(1.0)
        causal = f64(dr["u"]) < p_causal
        bt = f64(dr["b"]) * math.sqrt(sb2) * causal
        eps = (math.sqrt(1 - alpha) * f64(dr["e1"]) + math.sqrt(alpha) * dinv * (Z @ f64(dr["e2"])) / math.sqrt(k)) / math.sqrt(n)
        betas.append((R @ bt + eps))
        trues.append(bt)
        if ld_dtype == "int8":
            Rq = torch.round(R * 127.0)
        elif ld_dtype == "int16":
            Rq = torch.round(R * 32767.0)
        else:
            Rq = R
        if symmetric:
            data.append(Rq.reshape(-1).to(udt))
            lens.append(torch.full((B,), B, dtype=torch.int64, device=dev))
            lbs.append(torch.full((B,), row0, dtype=torch.int32, device=dev))
        else:
            mask = torch.ones(B, B, dtype=torch.bool, device=dev).triu_(1)
            data.append(Rq[mask].to(udt))            # row-major order == CSR data order
            lens.append(torch.arange(B - 1, -1, -1, dtype=torch.int64, device=dev))
            lbs.append(torch.arange(row0 + 1, row0 + B + 1, dtype=torch.int32, device=dev))
        row0 += B
    lens = torch.cat(lens) if lens else torch.zeros(0, dtype=torch.int64, device=dev)
    indptr = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(lens, 0)
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt, device=dev)
    return {
        "ld_data": cat(data, udt),
        "ld_indptr": indptr,
        "ld_left_bound": cat(lbs, torch.int32),
        "std_beta": cat(betas, torch.float64).to(float_dtype),
        "n_per_snp": torch.full((M,), float(n), dtype=torch.float64, device=dev),
        "beta_true": cat(trues, torch.float64),
        "dq_scale": dequantize_scale(ld_dtype),
        "block_sizes": list(block_sizes),
    }


def lognormal_sizes(M, median=650, sigma=0.6, lo=64, hi=4096, seed=SEED):
    """LDetect-like genome-wide block sizes: log-normal around `median` SNPs (EUR LDetect has ~1,700 blocks for the
    ~1.1 M HapMap3 SNPs), clipped to [lo, hi], summing to M."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 2 ** 40]))
    sizes, tot = [], 0
    while tot < M:
        s = int(np.clip(np.rint(np.exp(np.log(median) + sigma * rng.standard_normal())), lo, hi))
        s = min(s, M - tot)
        sizes.append(s)
        tot += s
    return sizes


def e_step_inputs(std_beta, n_per_snp, pi, sigma_epsilon, tau_beta, lambda_min=0.0, float_dtype=torch.float32):
    """
    The per-SNP vectors VIPRS.e_step() hands to cpp_e_step (VIPRS.py:400-406,418), computed in float64
    and cast like the reference: (u_logs, sqrt_half_var_tau, mu_mult, var_tau).
    """
    n = n_per_snp.to(torch.float64)
    var_tau = n * (1.0 + lambda_min) / sigma_epsilon + tau_beta
    mu_mult = (n / (var_tau * sigma_epsilon)).to(float_dtype)
    u_logs = (math.log(pi) - math.log(1.0 - pi) + 0.5 * (math.log(tau_beta) - torch.log(var_tau))).to(float_dtype)
    shvt = torch.sqrt(0.5 * var_tau).to(float_dtype)
    return u_logs, shvt, mu_mult, var_tau
