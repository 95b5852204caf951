"""
Parity at BASELINE.json's full sizes (C2, C4 and C3 shapes: 1,101,824 SNPs = 269 LD blocks x 4096).

LD blocks are independent inside one sweep (e_step.hpp:389-392: row j only touches columns of its own block), so every
block of a genome-wide sweep must equal the oracle run on that block alone with the same inputs -- an exact parity
check at the full size that costs the oracle only a few blocks.  Two sweeps are run so that the second one starts
from a non-trivial state.  Size-independent property checked next to it: q == dq (R - I) eta for the sampled blocks.
"""
import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu

N_BLOCKS, BLOCK = 269, 4096
SAMPLE = (0, 137, 268)


@pytest.fixture(scope="module")
def vb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import viprs_b200
    return viprs_b200


def _inputs(ld_dtype):
    import torch
    from viprs_b200 import synth
    sizes = synth.block_sizes_for(N_BLOCKS * BLOCK, BLOCK)
    assert len(sizes) == N_BLOCKS and all(s == BLOCK for s in sizes)
    inp = synth.make_inputs(sizes, ld_dtype=ld_dtype, float_dtype=torch.float32, device="cuda")
    return inp, int(sum(sizes))


def _block_host(inp, b):
    """magenpy-layout arrays of block b alone (row / column indices rebased to the block)."""
    r0, r1 = b * BLOCK, (b + 1) * BLOCK
    ip = inp["ld_indptr"][r0:r1 + 1].cpu().numpy()
    data = inp["ld_data"][int(ip[0]):int(ip[-1])].cpu().numpy()
    lb = (inp["ld_left_bound"][r0:r1].cpu().numpy() - r0).astype(np.int32)
    return lb, (ip - ip[0]).astype(np.int64), data, slice(r0, r1)


def _dense(lb, ip, data):
    B = lb.shape[0]
    R = np.zeros((B, B))
    for j in range(B):
        s, e = ip[j], ip[j + 1]
        R[j, lb[j]:lb[j] + (e - s)] = data[s:e]
    return R + R.T


def test_c2_full_size_blocks_match_oracle(vb, oracle_built):
    import torch
    from viprs_b200 import synth
    inp, M = _inputs("int8")
    pi, se = 0.01, 0.8
    ul, sv, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, pi * M / (1 - se))
    ld = vb.DeviceLD(inp["ld_data"], inp["ld_indptr"], inp["ld_left_bound"])
    assert ld.n_blocks == N_BLOCKS and ld.max_block == BLOCK
    st = {k: torch.zeros(M, dtype=torch.float32, device="cuda") for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = torch.full((M,), pi, dtype=torch.float32, device="cuda")
    for _ in range(2):
        vb.e_step_device(ld, inp["std_beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                         ul, sv, mm, inp["dq_scale"], True)
    torch.cuda.synchronize()
    T = np.float32
    for b in SAMPLE:
        lb, ip, data, sl = _block_host(inp, b)
        h = lambda t: t[sl].cpu().numpy().copy()
        ref = {k: np.zeros(BLOCK, T) for k in ("var_mu", "eta", "q", "eta_diff")}
        ref["var_gamma"] = np.full(BLOCK, pi, T)
        for _ in range(2):
            oracle_built.e_step(lb, ip, data, h(inp["std_beta"]), ref["var_gamma"], ref["var_mu"], ref["eta"], ref["q"],
                                ref["eta_diff"], h(ul), h(sv), h(mm), inp["dq_scale"], 1, True)
        for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
            assert relmax(h(st[k]), ref[k]) <= 1e-4, (b, k, relmax(h(st[k]), ref[k]))
        eta = h(st["eta"]).astype(np.float64)
        assert relmax(h(st["q"]), inp["dq_scale"] * (_dense(lb, ip, data.astype(np.float64)) @ eta)) <= 1e-4, b
    ld.destroy()


def test_c4_full_size_blocks_match_oracle(vb, oracle_built):
    import torch
    inp, M = _inputs("int16")
    K = 4
    d = 2.0 ** np.linspace(-3, 0, K)                                      # VIPRSMix.py:52
    pis, se = 0.01 * np.ones(K) / K, 0.8
    tau = d * (M * np.dot(1. / d, pis) / (1 - se))                        # VIPRSMix.py:155-161
    n = inp["n_per_snp"].cpu().numpy()
    vt = n[:, None] / se + tau
    T = np.float32
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(T))).cuda()
    ul = c(np.log(pis) - np.log1p(-pis) + .5 * (np.log(tau) - np.log(vt)))
    sv, mm = c(np.sqrt(.5 * vt)), c(n[:, None] / (vt * se))
    lnp = torch.full((M,), float(np.log(1 - pis.sum())), dtype=torch.float32, device="cuda")
    ld = vb.DeviceLD(inp["ld_data"], inp["ld_indptr"], inp["ld_left_bound"])
    st = {"var_gamma": c(np.tile(pis, (M, 1))), "var_mu": torch.zeros(M, K, dtype=torch.float32, device="cuda")}
    for k in ("eta", "q", "eta_diff"):
        st[k] = torch.zeros(M, dtype=torch.float32, device="cuda")
    for _ in range(2):
        vb.e_step_mixture_device(ld, inp["std_beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                 lnp, ul, sv, mm, inp["dq_scale"], True)
    torch.cuda.synchronize()
    for b in SAMPLE[:2]:
        lb, ip, data, sl = _block_host(inp, b)
        h = lambda t: np.ascontiguousarray(t[sl].cpu().numpy())
        ref = {"var_gamma": np.ascontiguousarray(np.tile(pis, (BLOCK, 1)).astype(T)), "var_mu": np.zeros((BLOCK, K), T),
               "eta": np.zeros(BLOCK, T), "q": np.zeros(BLOCK, T), "eta_diff": np.zeros(BLOCK, T)}
        for _ in range(2):
            oracle_built.e_step_mixture(lb, ip, data, h(inp["std_beta"]), ref["var_gamma"], ref["var_mu"], ref["eta"],
                                        ref["q"], ref["eta_diff"], h(lnp), h(ul), h(sv), h(mm), inp["dq_scale"], 1, True)
        for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
            assert relmax(h(st[k]), ref[k]) <= 1e-4, (b, k, relmax(h(st[k]), ref[k]))
    ld.destroy()


def test_c3_full_size_grid_columns_match_oracle(vb, oracle_built):
    """256 grid columns over the genome-wide LD; the oracle re-runs three of them on one block (grid columns are
    independent of each other, e_step.hpp:606-610)."""
    import torch
    from scipy.stats import norm
    inp, M = _inputs("int8")
    G = 256
    pis = np.logspace(np.log10(max(10. / M, 1e-5)), np.log10(min(1e4 / M, 0.2)), 16)      # HyperparameterGrid.py:193-205
    p0 = max(0.1, norm.cdf((1e-5 - 0.1) / 0.1))
    ses = 1. - norm.ppf(np.linspace(p0, 0.9, 16), 0.1, 0.1)                                # :146-163
    pi_g = np.tile(pis, 16)
    se_g = np.repeat(ses, 16)                                                              # sigma_epsilon-major (:238-245)
    tau_g = pi_g * M / (1 - se_g)
    T = np.float32
    n = float(inp["n_per_snp"][0].item())
    vt = n / se_g + tau_g                                                                  # n_per_snp is constant here
    row = lambda v: torch.from_numpy(np.ascontiguousarray(v.astype(T))).cuda()[:, None].expand(G, M).contiguous().t()
    ul = row(np.log(pi_g) - np.log1p(-pi_g) + .5 * (np.log(tau_g) - np.log(vt)))
    hv, mm = row(.5 * vt), row(n / (vt * se_g))
    ld = vb.DeviceLD(inp["ld_data"], inp["ld_indptr"], inp["ld_left_bound"])
    z = lambda: torch.zeros(G, M, dtype=torch.float32, device="cuda").t()
    st = {k: z() for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = row(pi_g)
    act = torch.arange(G, dtype=torch.int32, device="cuda")
    for _ in range(2):
        vb.e_step_grid_device(ld, inp["std_beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                              ul, hv, mm, inp["dq_scale"], act)
    torch.cuda.synchronize()
    cols = np.array([0, 100, 255], dtype=np.int32)
    b = 137
    lb, ip, data, sl = _block_host(inp, b)
    F = lambda t: np.asfortranarray(t[sl][:, cols.tolist()].cpu().numpy())
    ref = {k: np.zeros((BLOCK, len(cols)), T, order="F") for k in ("var_mu", "eta", "q", "eta_diff")}
    ref["var_gamma"] = np.asfortranarray(np.tile(pi_g[cols], (BLOCK, 1)).astype(T))
    for _ in range(2):
        oracle_built.e_step_grid(lb, ip, data, inp["std_beta"][sl].cpu().numpy(), ref["var_gamma"], ref["var_mu"], ref["eta"],
                                 ref["q"], ref["eta_diff"], F(ul), F(hv), F(mm), inp["dq_scale"],
                                 np.arange(len(cols), dtype=np.int32), 1, True)
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(F(st[k]), ref[k]) <= 1e-4, (k, relmax(F(st[k]), ref[k]))
    ld.destroy()
