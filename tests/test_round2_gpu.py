"""
GPU parity tests added in round 2 (all through the C ABI, against the compiled reference / the oracle):

  * tiled sweep: LD blocks larger than 4096 SNPs (float64 10,240-SNP blocks = BASELINE configs[4] shape, int8
    9,000-SNP block) and banded LD where a whole "chromosome" is one block;
  * q is in/out like the reference's: q_in that eta_in does not explain is carried through the sweeps
    (`param_0` warm start: eta != 0 next to q = 0, VIPRS.py:339-357);
  * BASELINE hyper-parameters (n = 3e5, h2 = 0.3, pi = 0.01, sigma_epsilon = 0.8) on one 4096-SNP C2 block (int8)
    and one C4 block (int16, K = 4) after {10, 50} sweeps, against the float32 AND the float64 reference; the
    measured numbers go to gpurun_out/parity_floor.json;
  * 10-iteration EM history (ELBO, pi, sigma_epsilon, tau_beta) on a 16-block slice of the C2 workload.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, relmax
from tests_util import make_block_ld

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import viprs_b200
    return viprs_b200


def _hyper(rng, M, T, pi=0.02, se=0.8):
    n = np.floor(rng.uniform(4e4, 6e4, M))
    tau = pi * M / (1 - se)
    vt = n / se + tau
    u_logs = (np.log(pi) - np.log(1 - pi) + .5 * (np.log(tau) - np.log(vt))).astype(T)
    return u_logs, np.sqrt(.5 * vt).astype(T), (n / (vt * se)).astype(T), pi


def _sweeps(fn, P, T, hy, n_sweeps, st=None):
    M = P["M"]
    u_logs, shvt, mm, pi = hy
    if st is None:
        st = {k: np.zeros(M, T) for k in ("var_mu", "eta", "q", "eta_diff")}
        st["var_gamma"] = np.full(M, pi, T)
    for _ in range(n_sweeps):
        fn(P["lb"], P["indptr"], P["data"], P["beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"],
           st["eta_diff"], u_logs, shvt, mm, P["dq"], 1, True)
    return st


def _device_sweeps(vb, P, T, hy, n_sweeps):
    """The ONE-PASS sweep through the device entry (viprs_b200_e_step_f32 / _f64): what fit() runs every iteration.  The
    float32 host drop-ins (vb.cpp_e_step) take the incremental route, which only exists as kernel version 1."""
    import torch
    M = P["M"]
    u_logs, shvt, mm, pi = hy
    td = torch.float32 if T == np.float32 else torch.float64
    dev = {k: torch.zeros(M, dtype=td, device="cuda") for k in ("var_mu", "eta", "q", "eta_diff")}
    dev["var_gamma"] = torch.full((M,), pi, dtype=td, device="cuda")
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    beta, ul, sv, mmd = c(P["beta"]), c(u_logs), c(shvt), c(mm)
    for _ in range(n_sweeps):
        vb.e_step_device(ld, beta, dev["var_gamma"], dev["var_mu"], dev["eta"], dev["q"], dev["eta_diff"], ul, sv, mmd, P["dq"], True)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in dev.items()}
    ld.destroy()
    return out


def _record(name, payload):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_floor.json")
    d = {}
    if os.path.exists(path):
        try:
            d = json.load(open(path))
        except Exception:
            d = {}
    d[name] = payload
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


# ---------------------------------------------------------------------------------------------------------
# tiled sweep
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tn,un,blocks", [("f64", "f64", (5000, 300)), ("f32", "i8", (9000, 64, 4500)),
                                          ("f64", "i16", (4097,)), ("f32", "f32", (6200,))])
def test_tiled_blocks_match_oracle(vb, oracle_built, tn, un, blocks):
    T = np.float32 if tn == "f32" else np.float64
    U = {"i8": np.int8, "i16": np.int16, "f32": np.float32, "f64": np.float64}[un]
    rng = np.random.default_rng(len(blocks) * 31 + blocks[0])
    P = make_block_ld(rng, blocks, U, T)
    hy = _hyper(rng, P["M"], T)
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    assert ld.n_blocks == len(blocks) and ld.max_block == max(blocks)
    assert ld.n_phases == -(-max(blocks) // 1024) and ld.n_units > ld.n_blocks and ld.ext_elems > 0
    ld.destroy()
    ref = _sweeps(oracle_built.e_step, P, T, hy, 3)
    got = _sweeps(vb.cpp_e_step, P, T, hy, 3)
    tol = 1e-4 if T == np.float32 else 1e-10
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(got[k], ref[k]) <= tol, (k, relmax(got[k], ref[k]))


def test_c5_shape_block_float64(vb, oracle_built):
    """BASELINE configs[4] shape: float64 state + float64 LD, one 10,240-SNP block (~5k stored entries per row)."""
    T = np.float64
    rng = np.random.default_rng(55)
    P = make_block_ld(rng, (10240, 128), np.float64, T)
    hy = _hyper(rng, P["M"], T)
    ref = _sweeps(oracle_built.e_step, P, T, hy, 2)
    got = _sweeps(vb.cpp_e_step, P, T, hy, 2)
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(got[k], ref[k]) <= 1e-10, (k, relmax(got[k], ref[k]))


def _banded(rng, M, w, U, T):
    """Windowed LD: row j stores columns j+1 .. min(j+w, M-1) -- no independent blocks at all."""
    lens = np.minimum(w, M - 1 - np.arange(M)).astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    dist = np.concatenate([np.arange(1, k + 1) for k in lens])
    vals = 0.35 * np.exp(-dist / (0.2 * w)) * np.cos(dist * 0.37) + 0.01 * rng.standard_normal(dist.shape[0])
    if U == np.int8:
        data, dq = np.rint(vals * 127).astype(np.int8), 1. / 127
    else:
        data, dq = vals.astype(U), 1.
    beta = (rng.standard_normal(M) / np.sqrt(5e4) + 0.004 * (rng.random(M) < 0.02) * rng.standard_normal(M)).astype(T)
    return {"M": M, "data": data, "indptr": indptr, "lb": np.arange(1, M + 1, dtype=np.int32), "beta": beta, "dq": dq}


@pytest.mark.parametrize("tn,un", [("f32", "i8"), ("f64", "f64")])
def test_banded_ld_is_swept_in_order(vb, oracle_built, tn, un):
    T = np.float32 if tn == "f32" else np.float64
    U = np.int8 if un == "i8" else np.float64
    rng = np.random.default_rng(77)
    P = _banded(rng, 7001, 650, U, T)
    hy = _hyper(rng, P["M"], T)
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    assert ld.n_blocks == 1 and ld.n_phases == -(-7001 // 1024)
    ld.destroy()
    ref = _sweeps(oracle_built.e_step, P, T, hy, 3)
    got = _sweeps(vb.cpp_e_step, P, T, hy, 3)
    tol = 1e-4 if T == np.float32 else 1e-10
    floor = {k: 0.0 for k in ref}
    if T == np.float32:
        # var_mu / var_gamma keep their OLD value when |eta_diff| < eps (e_step.hpp:410-413): on this matrix the float32
        # reference already sits 2e-3 / 2e-2 from its own float64 run, so these two are held to that floor; eta and q,
        # which the skip branch leaves untouched by at most eps, are held to 1e-4 (float64 state: everything to 1e-10)
        P64 = dict(P, beta=P["beta"].astype(np.float64))
        hy64 = tuple(np.asarray(a, np.float64) if isinstance(a, np.ndarray) else a for a in hy)
        ref64 = _sweeps(oracle_built.e_step, P64, np.float64, hy64, 3)
        floor = {k: relmax(ref[k], ref64[k]) for k in ("var_gamma", "var_mu")}
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(got[k], ref[k]) <= max(tol, floor.get(k, 0.0)), (k, relmax(got[k], ref[k]), floor.get(k))


# ---------------------------------------------------------------------------------------------------------
# q is in/out
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tn,un,blocks", [("f32", "i8", (700, 33, 257)), ("f64", "f64", (300, 120)),
                                          ("f64", "f64", (4500,))])
@pytest.mark.parametrize("q_in", ["zero", "garbage"])
def test_incoming_q_is_honoured(vb, oracle_built, tn, un, blocks, q_in):
    """eta_in != 0 with q_in = 0 (a param_0 warm start, VIPRS.py:339-357) or an arbitrary q_in: the reference keeps
    q incrementally, so the unexplained part of q_in stays; the drop-in must reproduce that."""
    T = np.float32 if tn == "f32" else np.float64
    U = np.int8 if un == "i8" else np.float64
    rng = np.random.default_rng(5)
    P = make_block_ld(rng, blocks, U, T)
    M = P["M"]
    hy = _hyper(rng, M, T)

    def start():
        st = {"var_gamma": rng0.uniform(0.01, 0.3, M).astype(T), "var_mu": (0.01 * rng0.standard_normal(M)).astype(T),
              "eta_diff": np.zeros(M, T)}
        st["eta"] = st["var_gamma"] * st["var_mu"]
        st["q"] = np.zeros(M, T) if q_in == "zero" else (0.003 * rng0.standard_normal(M)).astype(T)
        return st

    rng0 = np.random.default_rng(9)
    ref = _sweeps(oracle_built.e_step, P, T, hy, 2, start())
    rng0 = np.random.default_rng(9)
    got = _sweeps(vb.cpp_e_step, P, T, hy, 2, start())
    tol = 1e-4 if T == np.float32 else 1e-10
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(got[k], ref[k]) <= tol, (k, relmax(got[k], ref[k]))


def test_fit_with_param_0_matches_oracle(vb, oracle_built):
    from oracle import cpu as ocpu
    from viprs_b200.model import VIPRS
    T = np.float32
    rng = np.random.default_rng(21)
    P = make_block_ld(rng, (300, 517, 64), np.int8, T)
    M = P["M"]
    n = np.full(M, 5e4)
    param_0 = {"mu": {1: (0.004 * rng.standard_normal(M)).astype(T)}, "gamma": {1: rng.uniform(0.01, 0.2, M).astype(T)}}
    theta_0 = {"pi": 0.02, "sigma_epsilon": 0.8}
    o = ocpu.OracleVIPRS({1: (P["data"], P["indptr"], P["lb"])}, {1: P["beta"]}, {1: n}, float_precision="float32",
                         dequantize_on_the_fly=True)
    o.run(4, dict(theta_0), param_0)
    m = VIPRS(data={1: dict(ld_data=P["data"], ld_indptr=P["indptr"], ld_left_bound=P["lb"], std_beta=P["beta"], n_per_snp=n)},
              float_precision="float32")
    m.fit(max_iter=4, min_iter=10, theta_0=dict(theta_0), param_0=param_0)
    got = np.array(m.history["ELBO"][1:])
    assert np.max(np.abs(got - np.array(o.history["ELBO"])) / np.abs(np.array(o.history["ELBO"]))) <= 1e-4
    assert relmax(m.post_mean_beta[1], o.eta[1]) <= 1e-4
    assert relmax(m.q[1].cpu().numpy(), o.q[1]) <= 1e-4


# ---------------------------------------------------------------------------------------------------------
# BASELINE hyper-parameters past 2 sweeps
# ---------------------------------------------------------------------------------------------------------
def _baseline_block(ld_dtype, seed=7209):
    import torch
    from viprs_b200 import synth
    inp = synth.make_inputs([4096], ld_dtype=ld_dtype, float_dtype=torch.float32, device="cpu", seed=seed)   # n = 3e5, h2 = 0.3
    return inp


@pytest.mark.parametrize("n_sweeps", [10, 50])
def test_c2_block_baseline_params(vb, oracle_built, n_sweeps):
    """int8 LD, one 4096-SNP block of the C2 workload, BASELINE's n / h2 / pi / sigma_epsilon, {10, 50} sweeps."""
    from viprs_b200 import synth
    inp = _baseline_block("int8")
    M, pi, se = 4096, 0.01, 0.8
    res = {}
    for fdt, T in (("float32", np.float32), ("float64", np.float64)):
        import torch
        ul, sv, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, pi * 1101824 / (1 - se),
                                            float_dtype=getattr(torch, fdt))
        P = {"M": M, "lb": inp["ld_left_bound"].numpy(), "indptr": inp["ld_indptr"].numpy(), "data": inp["ld_data"].numpy(),
             "beta": inp["std_beta"].numpy().astype(T), "dq": inp["dq_scale"]}
        hy = (ul.numpy(), sv.numpy(), mm.numpy(), pi)
        res["ref_" + fdt] = _sweeps(oracle_built.e_step, P, T, hy, n_sweeps)
        if T == np.float32:
            res["got3"] = _sweeps(vb.cpp_e_step, P, T, hy, n_sweeps)
            os.environ["VIPRS_B200_LIMBS"] = "4"               # 28-bit fixed-point eta_old in the dp4a backward dots
            try:
                res["got4"] = _sweeps(vb.cpp_e_step, P, T, hy, n_sweeps)
            finally:
                del os.environ["VIPRS_B200_LIMBS"]
            # the incremental-q route (the reference's own q bookkeeping; what the host-state round trip and `e2e` run)
            ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
            res["incr"] = _sweeps(lambda lb, ip, dat, *a: vb.cpp_e_step_resident(ld, *a[:-3], a[-3], False), P, T, hy, n_sweeps)

            # the ONE-PASS sweep (what fit() runs every iteration and bench.py times: IDP.4A backward dots on a block-scaled
            # fixed-point eta_old) through the device entry, with 3 (21-bit) and 4 (28-bit) digits: the host drop-ins above
            # take the incremental route for float32, which has no backward dots at all
            def onepass(lb, ip, dat, beta, g, mu, eta, q, diff, ul_, sv_, mm_, dq, *rest):
                dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (beta, g, mu, eta, q, diff, ul_, sv_, mm_)]
                vb.e_step_device(ld, *dev, dq, True)
                torch.cuda.synchronize()
                for host, d_ in zip((g, mu, eta, q, diff), (dev[1], dev[2], dev[3], dev[4], dev[5])):
                    host[...] = d_.cpu().numpy()
            res["one3"] = _sweeps(onepass, P, T, hy, n_sweeps)
            os.environ["VIPRS_B200_LIMBS"] = "4"
            try:
                res["one4"] = _sweeps(onepass, P, T, hy, n_sweeps)
            finally:
                del os.environ["VIPRS_B200_LIMBS"]
            ld.destroy()
    rec = {}
    for k in ("eta", "var_gamma", "var_mu", "q"):
        rec[k] = {"floor_ref32_vs_ref64": relmax(res["ref_float32"][k], res["ref_float64"][k]),
                  "ours_vs_ref32": relmax(res["got3"][k], res["ref_float32"][k]),
                  "ours_vs_ref64": relmax(res["got3"][k], res["ref_float64"][k]),
                  "ours_limbs4_vs_ref64": relmax(res["got4"][k], res["ref_float64"][k]),
                  "limbs3_vs_limbs4": relmax(res["got3"][k], res["got4"][k]),
                  "incremental_vs_ref32": relmax(res["incr"][k], res["ref_float32"][k]),
                  "incremental_vs_ref64": relmax(res["incr"][k], res["ref_float64"][k]),
                  "onepass_vs_ref32": relmax(res["one3"][k], res["ref_float32"][k]),
                  "onepass_vs_ref64": relmax(res["one3"][k], res["ref_float64"][k]),
                  "onepass_limbs4_vs_ref64": relmax(res["one4"][k], res["ref_float64"][k]),
                  "onepass_limbs3_vs_limbs4": relmax(res["one3"][k], res["one4"][k])}
    _record(f"c2_block_int8_{n_sweeps}_sweeps", rec)
    print(json.dumps(rec))
    for k in ("eta", "var_gamma"):
        r = rec[k]
        # north star: within 1e-4 of the reference; where the float32 reference itself sits further than that from its
        # float64 self, "the reference" is only defined to that floor: require to be as close to ref64 as ref32 is (x2)
        assert r["ours_vs_ref32"] <= 1e-4 or r["ours_vs_ref64"] <= 2 * r["floor_ref32_vs_ref64"], (k, r)
        # the 21-bit fixed-point representation adds nothing visible next to float32 rounding
        assert r["limbs3_vs_limbs4"] <= max(1e-4, 2 * r["floor_ref32_vs_ref64"]), (k, r)
        assert r["incremental_vs_ref32"] <= 1e-4 or r["incremental_vs_ref64"] <= 2 * r["floor_ref32_vs_ref64"], (k, r)
        assert r["onepass_vs_ref32"] <= 1e-4 or r["onepass_vs_ref64"] <= 2 * r["floor_ref32_vs_ref64"], (k, r)
        assert r["onepass_limbs3_vs_limbs4"] <= max(1e-4, 2 * r["floor_ref32_vs_ref64"]), (k, r)


@pytest.mark.parametrize("n_sweeps", [10, 50])
def test_c4_block_baseline_params(vb, oracle_built, n_sweeps):
    """int16 LD, K = 4 mixture, one 4096-SNP block of the C4 workload, BASELINE parameters."""
    inp = _baseline_block("int16")
    M, K, se = 4096, 4, 0.8
    d = 2.0 ** np.linspace(-3, 0, K)
    pis = 0.01 * np.ones(K) / K
    tau = d * (1101824 * np.dot(1. / d, pis) / (1 - se))
    n = inp["n_per_snp"].numpy()
    vt = n[:, None] / se + tau
    res = {}
    for T in (np.float32, np.float64):
        C = lambda a: np.ascontiguousarray(a.astype(T))
        ul, sv, mm = C(np.log(pis) - np.log1p(-pis) + .5 * (np.log(tau) - np.log(vt))), C(np.sqrt(.5 * vt)), C(n[:, None] / (vt * se))
        lnp = np.full(M, np.log(1 - pis.sum()), T)
        for name, fn in (("ref", oracle_built.e_step_mixture),) + ((("got", vb.cpp_e_step_mixture),) if T == np.float32 else ()):
            st = {"var_gamma": C(np.tile(pis, (M, 1))), "var_mu": np.zeros((M, K), T), "eta": np.zeros(M, T),
                  "q": np.zeros(M, T), "eta_diff": np.zeros(M, T)}
            for _ in range(n_sweeps):
                fn(inp["ld_left_bound"].numpy(), inp["ld_indptr"].numpy(), inp["ld_data"].numpy(),
                   inp["std_beta"].numpy().astype(T), st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                   lnp, ul, sv, mm, inp["dq_scale"], 1, True)
            res[name + ("32" if T == np.float32 else "64")] = st
        if T == np.float32:
            # the one-pass mixture sweep (what VIPRSMix.fit runs every iteration) through the device entry: the float32 host
            # drop-in above takes the incremental route
            import torch
            ld = vb.DeviceLD(inp["ld_data"].numpy(), inp["ld_indptr"].numpy(), inp["ld_left_bound"].numpy())
            dev = {"beta": torch.from_numpy(inp["std_beta"].numpy().astype(T)).cuda(), "lnp": torch.from_numpy(lnp).cuda(),
                   "ul": torch.from_numpy(ul).cuda(), "sv": torch.from_numpy(sv).cuda(), "mm": torch.from_numpy(mm).cuda(),
                   "var_gamma": torch.from_numpy(C(np.tile(pis, (M, 1)))).cuda(), "var_mu": torch.zeros((M, K), dtype=torch.float32, device="cuda"),
                   "eta": torch.zeros(M, dtype=torch.float32, device="cuda"), "q": torch.zeros(M, dtype=torch.float32, device="cuda"),
                   "eta_diff": torch.zeros(M, dtype=torch.float32, device="cuda")}
            for _ in range(n_sweeps):
                vb.e_step_mixture_device(ld, dev["beta"], dev["var_gamma"], dev["var_mu"], dev["eta"], dev["q"], dev["eta_diff"],
                                         dev["lnp"], dev["ul"], dev["sv"], dev["mm"], inp["dq_scale"], True)
            torch.cuda.synchronize()
            res["one32"] = {k: dev[k].cpu().numpy() for k in ("var_gamma", "var_mu", "eta", "q", "eta_diff")}
            ld.destroy()
    rec = {}
    for k in ("eta", "var_gamma", "var_mu", "q"):
        rec[k] = {"floor_ref32_vs_ref64": relmax(res["ref32"][k], res["ref64"][k]),
                  "ours_vs_ref32": relmax(res["got32"][k], res["ref32"][k]),
                  "ours_vs_ref64": relmax(res["got32"][k], res["ref64"][k]),
                  "onepass_vs_ref32": relmax(res["one32"][k], res["ref32"][k]),
                  "onepass_vs_ref64": relmax(res["one32"][k], res["ref64"][k])}
    _record(f"c4_block_int16_K4_{n_sweeps}_sweeps", rec)
    print(json.dumps(rec))
    for k in ("eta", "var_gamma"):
        r = rec[k]
        assert r["ours_vs_ref32"] <= 1e-4 or r["ours_vs_ref64"] <= 2 * r["floor_ref32_vs_ref64"], (k, r)
        assert r["onepass_vs_ref32"] <= 1e-4 or r["onepass_vs_ref64"] <= 2 * r["floor_ref32_vs_ref64"], (k, r)


# ---------------------------------------------------------------------------------------------------------
# EM history at BASELINE block size
# ---------------------------------------------------------------------------------------------------------
def test_em_history_c2_slice(vb, oracle_built):
    """10 EM iterations (E-step, M-step, ELBO) on a 16-block slice of the C2 workload (65,536 SNPs, int8 LD, n = 3e5):
    ELBO / pi / sigma_epsilon / tau_beta per iteration against the numpy + compiled-reference restatement of VIPRS.fit."""
    import torch
    from oracle import cpu as ocpu
    from viprs_b200 import synth
    from viprs_b200.model import VIPRS
    sizes = [4096] * 16
    inp = synth.make_inputs(sizes, ld_dtype="int8", float_dtype=torch.float32, device="cpu")
    theta_0 = {"pi": 0.01, "sigma_epsilon": 0.8}
    lb, ip, ldd = inp["ld_left_bound"].numpy(), inp["ld_indptr"].numpy(), inp["ld_data"].numpy()
    beta, n = inp["std_beta"].numpy(), inp["n_per_snp"].numpy()
    o = ocpu.OracleVIPRS({1: (ldd, ip, lb)}, {1: beta}, {1: n}, float_precision="float32", dequantize_on_the_fly=True)
    o.run(10, dict(theta_0))
    m = VIPRS(data={1: dict(ld_data=inp["ld_data"], ld_indptr=inp["ld_indptr"], ld_left_bound=inp["ld_left_bound"],
                            std_beta=inp["std_beta"], n_per_snp=inp["n_per_snp"])}, float_precision="float32",
              tracked_params=["pi", "sigma_epsilon", "tau_beta"])
    m.fit(max_iter=10, min_iter=20, theta_0=dict(theta_0), f_abs_tol=0., x_abs_tol=0.)
    rec = {}
    for key in ("ELBO", "pi", "sigma_epsilon", "tau_beta"):
        a = np.array(m.history[key][1:], dtype=np.float64)[:10]          # entry 0 is the state before the first iteration
        b = np.array(o.history[key], dtype=np.float64)
        rec[key] = float(np.max(np.abs(a - b) / np.abs(b)))
    _record("em_history_c2_16_blocks_10_iterations", rec)
    print(json.dumps(rec))
    for key, v in rec.items():
        assert v <= 1e-4, (key, v)
    assert relmax(m.post_mean_beta[1], o.eta[1]) <= 2e-4


# ---------------------------------------------------------------------------------------------------------
# device-resident EM iterations (scalar M-step / ELBO on the device, CUDA graph)
# ---------------------------------------------------------------------------------------------------------
def _small_data(seed=3, ld_dtype=np.int8):
    T = np.float32
    rng = np.random.default_rng(seed)
    P = make_block_ld(rng, (300, 517, 64, 1200), ld_dtype, T)
    n = np.floor(rng.uniform(4e4, 6e4, P["M"]))
    half = 300 + 517
    ip = P["indptr"]
    # two "chromosomes" so that pi = mean of per-chromosome means is exercised
    c1 = dict(ld_data=P["data"][:ip[half]], ld_indptr=ip[:half + 1], ld_left_bound=P["lb"][:half], std_beta=P["beta"][:half],
              n_per_snp=n[:half])
    c2 = dict(ld_data=P["data"][ip[half]:], ld_indptr=ip[half:] - ip[half], ld_left_bound=P["lb"][half:] - half,
              std_beta=P["beta"][half:], n_per_snp=n[half:])
    return {1: c1, 2: c2}


@pytest.mark.parametrize("kind", ["viprs", "mix"])
@pytest.mark.parametrize("graph_chunk", [4, 1])
def test_device_loop_matches_host_loop(vb, kind, graph_chunk):
    from viprs_b200.model import VIPRS, VIPRSMix
    data = _small_data()
    mk = (lambda: VIPRS(data=data, float_precision="float32", tracked_params=["pi", "sigma_epsilon", "tau_beta"])) if kind == "viprs" \
        else (lambda: VIPRSMix(data=data, K=3, float_precision="float32"))
    th = {"pi": 0.02, "sigma_epsilon": 0.8} if kind == "viprs" else {"pis": [0.01, 0.005, 0.005], "sigma_epsilon": 0.8}
    a = mk().fit(max_iter=9, min_iter=20, theta_0=dict(th), f_abs_tol=0., x_abs_tol=0.)
    b = mk().fit(max_iter=9, min_iter=20, theta_0=dict(th), f_abs_tol=0., x_abs_tol=0., device_loop=True, check_every=graph_chunk)
    ea, eb = np.array(a.history["ELBO"]), np.array(b.history["ELBO"])
    assert ea.shape == eb.shape == (10,)
    assert np.max(np.abs(ea - eb) / np.abs(ea)) <= 1e-10
    if kind == "viprs":
        for k in ("pi", "sigma_epsilon", "tau_beta"):
            assert np.allclose(a.history[k], b.history[k], rtol=1e-10, atol=0)
    for c in (1, 2):
        assert relmax(b.post_mean_beta[c], a.post_mean_beta[c]) <= 1e-6
        assert relmax(b.pip[c], a.pip[c]) <= 1e-6
        assert relmax(b.post_var_beta[c], a.post_var_beta[c]) <= 1e-6
    assert b.optim_result.nit == a.optim_result.nit and b.optim_result.message == a.optim_result.message


def test_device_iterations_grid_match_host_steps(vb):
    from viprs_b200.model import VIPRSGrid
    data = _small_data(seed=8)
    grid = [{"pi": p, "sigma_epsilon": s} for s in (0.7, 0.9) for p in (0.005, 0.02, 0.1)]

    def mk():
        m = VIPRSGrid(data=data, grid=grid, float_precision="float32")
        m._batched = True
        m._init_grid_hyper({})
        m.initialize_variational_parameters()
        m._active = list(range(len(grid)))
        return m
    a, b = mk(), mk()
    ref = []
    for _ in range(5):
        a.e_step(); a.m_step()
        ref.append(np.stack([a.elbo(), a.mse(), a.max_eta_diff()], axis=1))
    hist = b.em_iterations(5, graph=True)
    got = hist[:, :, :3]
    assert got.shape == (5, len(grid), 3)
    assert np.max(np.abs(got - np.array(ref)) / np.maximum(np.abs(np.array(ref)), 1e-12)) <= 1e-9
    assert np.allclose(b._hyp.tau_beta, a._hyp.tau_beta, rtol=1e-10)


@pytest.mark.parametrize("ld_dtype,blocks", [(np.int8, (300, 517, 64, 1200)), (np.float32, (257, 33)), (np.int8, (4600, 100))])
def test_fused_sums_match_the_sums_kernel(vb, ld_dtype, blocks, monkeypatch):
    """The reductions fused into the sweep's output role (viprs_b200_e_step_fused_f32, kernel version 2) against
    viprs_b200_sums_f32 run on the arrays the same sweep left behind -- including a tiled LD block and two chromosome
    segments."""
    import torch
    monkeypatch.setenv("VIPRS_B200_FAST", "2")
    from viprs_b200 import _lib
    from viprs_b200.ld import _stream_ptr
    from viprs_b200.model import VIPRS
    T = np.float32
    rng = np.random.default_rng(12)
    P = make_block_ld(rng, blocks, ld_dtype, T)
    n = np.floor(rng.uniform(4e4, 6e4, P["M"]))
    half = blocks[0]
    ip = P["indptr"]
    data = {1: dict(ld_data=P["data"][:ip[half]], ld_indptr=ip[:half + 1], ld_left_bound=P["lb"][:half], std_beta=P["beta"][:half],
                    n_per_snp=n[:half]),
            2: dict(ld_data=P["data"][ip[half]:], ld_indptr=ip[half:] - ip[half], ld_left_bound=P["lb"][half:] - half,
                    std_beta=P["beta"][half:], n_per_snp=n[half:])}
    m = VIPRS(data=data, float_precision="float32")
    m.initialize({"pi": 0.02, "sigma_epsilon": 0.8})
    for it in range(3):
        m.e_step()
        assert m._sums_fused
        fused = m._sums_dev.clone()
        L = _lib.lib()
        ref = torch.zeros_like(fused)
        rc = L.viprs_b200_sums_f32(m.M, 1, 0, 2, m._seg_dev.data_ptr(), m._g.data_ptr(), m._mu.data_ptr(), m._eta.data_ptr(),
                                   m._q.data_ptr(), m._diff.data_ptr(), m.std_beta_dev.data_ptr(), m.n_per_snp_dev.data_ptr(),
                                   m._theta_dev.data_ptr(), None, 2.0, m._ws.data_ptr(), m._ws.numel(), ref.data_ptr(), _stream_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        a, b = fused.cpu().numpy(), ref.cpu().numpy()
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) <= 1e-11, (it, a, b)
        m.m_step()


@pytest.mark.parametrize("tn,un,blocks", [("f32", "i8", (257, 64, 1, 2, 33, 700, 17)), ("f32", "i16", (4096, 500)),
                                          ("f32", "f32", (6200,)), ("f32", "i8", (4096,))])
def test_kernel_version_2_matches_oracle(vb, oracle_built, tn, un, blocks, monkeypatch):
    """The alternative chain / output-warp organisation of the register-resident kernel (VIPRS_B200_FAST=2)."""
    monkeypatch.setenv("VIPRS_B200_FAST", "2")
    T = np.float32
    U = {"i8": np.int8, "i16": np.int16, "f32": np.float32}[un]
    rng = np.random.default_rng(1000 + len(blocks))
    P = make_block_ld(rng, blocks, U, T)
    hy = _hyper(rng, P["M"], T)
    ref = _sweeps(oracle_built.e_step, P, T, hy, 3)
    got = _device_sweeps(vb, P, T, hy, 3)            # the device entry: the host drop-in would run version 1 (incremental route)
    # e_step.hpp:410-413: an update with |eta_diff| < eps is skipped and var_mu / var_gamma keep their previous value; a
    # SNP whose |eta_diff| sits at eps (1.2e-7) is skipped by one float32 evaluation order and not by another, so these
    # two arrays are compared where BOTH sides performed the update (eta / q move by at most eps either way)
    both = (ref["eta_diff"] != 0) & (got["eta_diff"] != 0)
    assert both.mean() > 0.5
    for k in ("eta", "q"):
        assert relmax(got[k], ref[k]) <= 1e-4, (k, relmax(got[k], ref[k]))
    for k in ("var_gamma", "var_mu", "eta_diff"):
        assert relmax(got[k][both], ref[k][both]) <= 1e-4, (k, relmax(got[k][both], ref[k][both]))


# ---------------------------------------------------------------------------------------------------------
# grid post-processing on the device (grid_utils.select_best_model / bayesian_model_average, pseudo-validation)
# ---------------------------------------------------------------------------------------------------------
def _fitted_grid(pathwise=False):
    from viprs_b200.model import VIPRSGrid
    data = _small_data(seed=8)
    grid = [{"pi": p, "sigma_epsilon": s} for s in (0.7, 0.9) for p in (0.005, 0.02, 0.1)]
    m = VIPRSGrid(data=data, grid=grid, float_precision="float32")
    m.fit(pathwise=pathwise, max_iter=6, min_iter=20, f_abs_tol=0., x_abs_tol=0.)
    return m, data, grid


@pytest.mark.parametrize("pathwise", [False, True])
def test_select_best_model_and_pseudo_validation(vb, pathwise):
    import torch
    m, data, grid = _fitted_grid(pathwise)
    G = len(grid)
    g = {c: m.var_gamma[c].cpu().numpy().astype(np.float64) for c in data}
    mu = {c: m.var_mu[c].cpu().numpy().astype(np.float64) for c in data}
    q = {c: m.q[c].cpu().numpy().astype(np.float64) for c in data}
    elbos = np.asarray(m._elbo_final).copy()
    rng = np.random.default_rng(0)
    vbeta = {c: np.asarray(data[c]["std_beta"], dtype=np.float64) + rng.standard_normal(len(data[c]["std_beta"])) / 300 for c in data}
    eta = np.concatenate([g[c] * mu[c] for c in data])
    qq = np.concatenate([q[c] for c in data])
    vb_all = np.concatenate([vbeta[c] for c in data])
    want = (eta * vb_all[:, None]).sum(0) ** 2 / (eta * (qq + eta)).sum(0)          # eval/pseudo_metrics.py:149-152
    got = m.pseudo_validate(vbeta)
    assert got.shape == (G,) and np.allclose(got, want, rtol=1e-6)
    best = int(np.argmax(want))
    m.select_best_model("pseudo_validation", vbeta)
    assert m.best_model_idx == best and m.n_models == 1
    for c in data:
        assert np.allclose(m.pip[c], g[c][:, best], rtol=1e-6) and np.allclose(m.post_mean_beta[c], (g[c] * mu[c])[:, best], rtol=1e-5)
    assert abs(m.pi - grid[best]["pi"]) < 1e-7 and abs(m.sigma_epsilon - grid[best]["sigma_epsilon"]) < 1e-6
    m2, _, _ = _fitted_grid(pathwise)
    m2.select_best_model("ELBO")
    assert m2.best_model_idx == int(np.argmax(elbos))


def test_bayesian_model_average(vb):
    m, data, grid = _fitted_grid(False)
    elbos = np.asarray(m._elbo_final).copy()
    w = np.exp(elbos - elbos.max()); w /= w.sum()                                  # scipy.special.softmax
    g = {c: m.var_gamma[c].cpu().numpy().astype(np.float64) for c in data}
    mu = {c: m.var_mu[c].cpu().numpy().astype(np.float64) for c in data}
    q = {c: m.q[c].cpu().numpy().astype(np.float64) for c in data}
    vt = {c: m.var_tau[c].cpu().numpy() for c in data}
    M = sum(len(data[c]["std_beta"]) for c in data)
    m.bayesian_model_average()
    ga = {c: (g[c] * w).sum(1) for c in data}
    mua = {c: (mu[c] * w).sum(1) for c in data}
    qa = {c: (q[c] * w).sum(1) for c in data}
    vta = {c: (vt[c] * w).sum(1) for c in data}
    zeta = {c: ga[c].astype(np.float32).astype(np.float64) * (mua[c].astype(np.float32).astype(np.float64) ** 2 + 1. / vta[c]) for c in data}
    for c in data:
        assert np.allclose(m.pip[c], ga[c], rtol=2e-6)
        assert np.allclose(m.post_mean_beta[c], ga[c] * mua[c], rtol=1e-5, atol=1e-12)
    pi = np.mean([ga[c].mean() for c in data])                                      # VIPRS.py:434 on the averaged state
    tau = pi * M / sum(zeta[c].sum() for c in data)
    assert abs(m.pi - pi) <= 1e-6 * pi and abs(m.tau_beta - tau) <= 1e-5 * tau and m.n_models == 1
    sg = sum((zeta[c] + qa[c] * (ga[c] * mua[c])).sum() for c in data)
    assert abs(m._sigma_g - sg) <= 1e-4 * abs(sg)


# ---------------------------------------------------------------------------------------------------------
# the reference's incremental q on the device (viprs_b200_e_step_*incremental_f32) and the host-state round trip
# that runs it in row chunks (viprs_b200_cpp_e_step*_resident)
# ---------------------------------------------------------------------------------------------------------
_INCR_BLOCKS = (700, 33, 257, 1, 2, 512, 300, 64, 129, 1000, 17, 400)      # 12 LD blocks -> 4 row chunks


def _incr_start(rng0, M, T, q_in, K=0):
    shape = (M, K) if K else (M,)
    st = {"var_gamma": np.ascontiguousarray(rng0.uniform(0.01, 0.3 / max(K, 1), shape).astype(T)),
          "var_mu": np.ascontiguousarray((0.01 * rng0.standard_normal(shape)).astype(T)), "eta_diff": np.zeros(M, T)}
    eta = st["var_gamma"] * st["var_mu"]
    st["eta"] = np.ascontiguousarray(eta.sum(axis=1) if K else eta).astype(T)
    st["q"] = np.zeros(M, T) if q_in == "zero" else (0.003 * rng0.standard_normal(M)).astype(T)
    return st


@pytest.mark.parametrize("un", ["i8", "i16", "f32"])
@pytest.mark.parametrize("q_in", ["zero", "garbage"])
def test_incremental_sweep_matches_oracle(vb, oracle_built, un, q_in):
    """cpp_e_step semantics with q in/out on device arrays: 3 sweeps from a warm start with an arbitrary q_in."""
    import torch
    T = np.float32
    U = {"i8": np.int8, "i16": np.int16, "f32": np.float32}[un]
    rng = np.random.default_rng(21)
    P = make_block_ld(rng, _INCR_BLOCKS, U, T)
    M = P["M"]
    hy = _hyper(rng, M, T)
    ref = _sweeps(oracle_built.e_step, P, T, hy, 3, _incr_start(np.random.default_rng(9), M, T, q_in))
    st = _incr_start(np.random.default_rng(9), M, T, q_in)
    u_logs, shvt, mm, _ = hy
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    dev = {k: torch.from_numpy(v).cuda() for k, v in st.items()}
    beta, ul, sv, m_ = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (P["beta"], u_logs, shvt, mm))
    for _ in range(3):
        vb.e_step_incremental_device(ld, beta, dev["var_gamma"], dev["var_mu"], dev["eta"], dev["q"], dev["eta_diff"],
                                     ul, sv, m_, P["dq"])
    torch.cuda.synchronize()
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(dev[k].cpu().numpy(), ref[k]) <= 1e-4, (k, relmax(dev[k].cpu().numpy(), ref[k]))
    ld.destroy()


def test_incremental_sweep_unsupported_for_float64(vb):
    """float64 state is not covered by the incremental sweep: the entry says so, it does not compute something else."""
    import torch
    rng = np.random.default_rng(3)
    P = make_block_ld(rng, (300, 120), np.float64, np.float64)
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    z = torch.zeros(P["M"], dtype=torch.float32, device="cuda")
    with pytest.raises(vb.ViprsB200Error) as ei:
        vb.e_step_incremental_device(ld, z, z.clone(), z.clone(), z.clone(), z.clone(), z.clone(), z, z, z, 1.0)
    assert ei.value.code == -6 or "unsupported" in str(ei.value).lower()
    ld.destroy()


_PINNED_KEEP = []


def _pinned(a, shift=0):
    """A page-locked (torch pin_memory) copy of `a` as a numpy array; shift > 0: a view that starts `shift` elements into
    the pinned allocation (its base is then not 16-byte aligned)."""
    import torch
    a = np.ascontiguousarray(a)
    buf = torch.empty(a.size + shift, dtype=torch.from_numpy(a).dtype).pin_memory()
    _PINNED_KEEP.append(buf)
    v = buf.numpy()[shift:shift + a.size].reshape(a.shape)
    v[...] = a
    return v


@pytest.mark.parametrize("tn,un", [("f32", "i8"), ("f32", "f32"), ("f64", "f64")])
@pytest.mark.parametrize("q_in,vouch", [("zero", False), ("garbage", False)])
@pytest.mark.parametrize("mem", ["pageable", "pinned", "pinned_shifted"])
def test_resident_host_state_matches_oracle(vb, oracle_built, tn, un, q_in, vouch, mem):
    """viprs_b200_cpp_e_step_resident: HOST state arrays on a resident LD, 3 calls.  float32: incremental sweep in row
    chunks on internal streams; float64: q offset + one-pass sweep.  Page-locked arrays (what bench.py's e2e leg passes)
    make the chunk copies truly asynchronous; chunk offsets that are not multiples of 16 bytes and array bases that are
    not 16-byte aligned must not matter."""
    if mem != "pageable" and (tn != "f32" or q_in == "zero"):
        pytest.skip("page-locked variants: the chunked float32 route, one start")
    T = np.float32 if tn == "f32" else np.float64
    U = {"i8": np.int8, "f32": np.float32, "f64": np.float64}[un]
    rng = np.random.default_rng(22)
    P = make_block_ld(rng, _INCR_BLOCKS, U, T)
    M = P["M"]
    hy = _hyper(rng, M, T)
    ref = _sweeps(oracle_built.e_step, P, T, hy, 3, _incr_start(np.random.default_rng(9), M, T, q_in))
    st = _incr_start(np.random.default_rng(9), M, T, q_in)
    u_logs, shvt, mm, _ = hy
    u_logs, shvt, mm = (np.ascontiguousarray(a) for a in (u_logs, shvt, mm))
    if mem != "pageable":
        st = {k: _pinned(v, 1 if (mem == "pinned_shifted" and k == "eta") else 0) for k, v in st.items()}
        P = dict(P, beta=_pinned(P["beta"]))
        u_logs, shvt, mm = _pinned(u_logs), _pinned(shvt, 3 if mem == "pinned_shifted" else 0), _pinned(mm)
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    for _ in range(3):
        vb.cpp_e_step_resident(ld, P["beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                               u_logs, shvt, mm, P["dq"], vouch)
    tol = 1e-4 if T == np.float32 else 1e-10
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(st[k], ref[k]) <= tol, (k, relmax(st[k], ref[k]))
    ld.destroy()


@pytest.mark.parametrize("K", [1, 3, 4])
@pytest.mark.parametrize("mem", ["pageable", "pinned"])
def test_resident_mixture_host_state_matches_oracle(vb, oracle_built, K, mem):
    """viprs_b200_cpp_e_step_mixture_resident (int16 LD, float32): chunked incremental mixture sweep, 3 calls, pageable
    and page-locked host arrays."""
    T = np.float32
    rng = np.random.default_rng(23)
    P = make_block_ld(rng, _INCR_BLOCKS, np.int16, T)
    M = P["M"]
    n = np.floor(rng.uniform(4e4, 6e4, M))[:, None]
    d = 2.0 ** np.linspace(-min(K - 1, 7), 0, K)
    pis = 0.03 * np.ones(K) / K
    tau = d * (M * np.dot(1. / d, pis) / 0.2)
    vt = n / 0.8 + tau
    u_logs = np.ascontiguousarray((np.log(pis) - np.log(1 - pis) + .5 * (np.log(tau) - np.log(vt))).astype(T))
    shvt, mm = np.ascontiguousarray(np.sqrt(.5 * vt).astype(T)), np.ascontiguousarray((n / (vt * 0.8)).astype(T))
    lnp = np.full(M, np.log(1 - pis.sum()), T)
    ref = _incr_start(np.random.default_rng(9), M, T, "garbage", K)
    got = _incr_start(np.random.default_rng(9), M, T, "garbage", K)
    if mem == "pinned":
        got = {k: _pinned(v) for k, v in got.items()}
        P = dict(P, beta=_pinned(P["beta"]))
        lnp, u_logs, shvt, mm = _pinned(lnp), _pinned(u_logs), _pinned(shvt), _pinned(mm)
    ld = vb.DeviceLD(P["data"], P["indptr"], P["lb"])
    for _ in range(3):
        oracle_built.e_step_mixture(P["lb"], P["indptr"], P["data"], P["beta"], ref["var_gamma"], ref["var_mu"], ref["eta"],
                                    ref["q"], ref["eta_diff"], lnp, u_logs, shvt, mm, P["dq"], 1, True)
        vb.cpp_e_step_mixture_resident(ld, P["beta"], got["var_gamma"], got["var_mu"], got["eta"], got["q"], got["eta_diff"],
                                       lnp, u_logs, shvt, mm, P["dq"], False)
    for k in ("eta", "var_gamma", "var_mu", "q", "eta_diff"):
        assert relmax(got[k], ref[k]) <= 1e-4, (k, relmax(got[k], ref[k]))
    ld.destroy()
