"""
CPU-only: pins the oracle (oracle/estep_port.c + the numpy restatement in oracle/cpu.py) to
 (a) the compiled, unmodified reference (oracle/_ref/libviprs_ref.so) -- bit for bit, and
 (b) the golden vectors produced by the reference's own Python classes (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from conftest import load_golden, relmax

import oracle
from oracle import cpu

LD_DT = {"i8": np.int8, "i16": np.int16, "f32": np.float32, "f64": np.float64}
COMBOS = [("f32", "i8"), ("f32", "i16"), ("f32", "f32"), ("f64", "i8"), ("f64", "i16"), ("f64", "f32"), ("f64", "f64")]


def _random_problem(rng, M, ld, T, blocks=(37, 64, 21), symmetric=False):
    """Random block LD in the reference layout + a random (non-initial) variational state."""
    from tests_util import make_block_ld
    return make_block_ld(rng, blocks, ld, T, symmetric)


def _state(rng, M, T, K=None, G=None):
    shape = (M,) if K is None and G is None else ((M, K) if K else (M, G))
    order = "F" if G else "C"
    g = np.asarray(rng.uniform(0.01, 0.6, shape).astype(T), order=order)
    mu = np.asarray((rng.standard_normal(shape) * 0.01).astype(T), order=order)
    return g, mu


def _per_snp(rng, M, T, cols=None):
    n = rng.uniform(4e4, 6e4, M)
    shape = (M,) if cols is None else (M, cols)
    pi = rng.uniform(0.005, 0.1, shape[1:] or None)
    se = rng.uniform(0.5, 0.9)
    tau = rng.uniform(50, 5000, shape[1:] or None)
    nn = n if cols is None else n[:, None]
    vt = nn / se + tau
    u_logs = (np.log(pi) - np.log(1 - pi) + .5 * (np.log(tau) - np.log(vt))).astype(T)
    return u_logs, vt, (nn / (vt * se)).astype(T)


@pytest.mark.parametrize("tn,un", COMBOS)
@pytest.mark.parametrize("low_memory", [True, False])
def test_port_matches_compiled_reference_e_step(oracle_built, tn, un, low_memory):
    if not oracle.have_ref():
        pytest.skip("compiled reference not present (built only where /root/reference exists)")
    T = np.float32 if tn == "f32" else np.float64
    rng = np.random.default_rng(11)
    from tests_util import make_block_ld
    P = make_block_ld(rng, (37, 64, 21, 1, 2), LD_DT[un], T, symmetric=not low_memory)
    M = P["M"]
    u_logs, vt, mm = _per_snp(rng, M, T)
    shvt = np.sqrt(.5 * vt).astype(T)
    outs = []
    for kind in ("reference", "port"):
        g, mu = _state(np.random.default_rng(5), M, T)
        eta = (g * mu).astype(T)
        q = np.zeros(M, T)
        diff = np.zeros(M, T)
        for _ in range(3):
            cpu.e_step(P["lb"], P["indptr"], P["data"], P["beta"], g, mu, eta, q, diff, u_logs, shvt, mm,
                       P["dq"], 1, low_memory, kind=kind)
        outs.append((g, mu, eta, q, diff))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("tn,un", [("f32", "i16"), ("f64", "f64"), ("f32", "f32")])
def test_port_matches_compiled_reference_mixture_and_grid(oracle_built, tn, un):
    if not oracle.have_ref():
        pytest.skip("compiled reference not present")
    T = np.float32 if tn == "f32" else np.float64
    rng = np.random.default_rng(12)
    from tests_util import make_block_ld
    P = make_block_ld(rng, (40, 33, 50), LD_DT[un], T)
    M, K, G = P["M"], 3, 4
    # mixture
    u_logs, vt, mm = _per_snp(rng, M, T, K)
    shvt = np.sqrt(.5 * vt).astype(T)
    lnp = np.full(M, np.log(0.8), T)
    res = []
    for kind in ("reference", "port"):
        g, mu = _state(np.random.default_rng(6), M, T, K=K)
        g = (g / 4).astype(T)
        eta = (g * mu).sum(axis=1).astype(T)
        q, diff = np.zeros(M, T), np.zeros(M, T)
        for _ in range(2):
            cpu.e_step_mixture(P["lb"], P["indptr"], P["data"], P["beta"], g, mu, eta, q, diff, lnp,
                               np.ascontiguousarray(u_logs), np.ascontiguousarray(shvt), np.ascontiguousarray(mm),
                               P["dq"], 1, True, kind=kind)
        res.append((g, mu, eta, q, diff))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    # grid
    u_logs, vt, mm = _per_snp(rng, M, T, G)
    F = np.asfortranarray
    res = []
    for kind in ("reference", "port"):
        g, mu = _state(np.random.default_rng(7), M, T, G=G)
        eta = F((g * mu).astype(T))
        q, diff = F(np.zeros((M, G), T)), F(np.zeros((M, G), T))
        for _ in range(2):
            cpu.e_step_grid(P["lb"], P["indptr"], P["data"], P["beta"], g, mu, eta, q, diff, F(u_logs),
                            F((.5 * vt).astype(T)), F(mm), P["dq"], np.array([0, 2, 3], np.int32), 1, True, kind=kind)
        res.append((g, mu, eta, q, diff))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_raw_sweeps_match_golden(oracle_built, kind):
    """cpp_e_step (f64 state, int16 LD) golden: two sweeps from the initial state."""
    if kind == "reference" and not oracle.have_ref():
        pytest.skip("compiled reference not present")
    d, ch = load_golden("cpp_e_step_f64_i16.npz")
    c = ch[0]
    M = len(c["std_beta"])
    st = {k: np.zeros(M) for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.full(M, 0.03)
    for sweep in (1, 2):
        cpu.e_step(c["ld_left_bound"], c["ld_indptr"], c["ld_data"], c["std_beta"], st["var_gamma"], st["var_mu"],
                   st["eta"], st["q"], st["eta_diff"], d["raw_u_logs"], d["raw_sqrt_half_var_tau"], d["raw_mu_mult"],
                   1. / 32767, 1, True, kind=kind)
        for k, a in st.items():
            assert np.array_equal(a, d[f"raw_sweep{sweep}_{k}"]), (sweep, k)


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_grid_sweeps_match_golden(oracle_built, kind):
    if kind == "reference" and not oracle.have_ref():
        pytest.skip("compiled reference not present")
    d, ch = load_golden("e_step_grid_f32_i8.npz")
    c = ch[0]
    M, G = d["grid_u_logs"].shape
    f32 = np.float32
    st = {k: np.zeros((M, G), f32, order="F") for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.asfortranarray(np.tile(d["grid_pis"].astype(f32), (M, 1)))
    for sweep in (1, 2, 3):
        cpu.e_step_grid(c["ld_left_bound"], c["ld_indptr"], c["ld_data"], c["std_beta"].astype(f32), st["var_gamma"],
                        st["var_mu"], st["eta"], st["q"], st["eta_diff"], np.asfortranarray(d["grid_u_logs"]),
                        np.asfortranarray(d["grid_half_var_tau"]), np.asfortranarray(d["grid_mu_mult"]),
                        f32(1. / 127), d["grid_active"], 1, True, kind=kind)
        if sweep in (1, 3):
            for k, a in st.items():
                assert np.array_equal(a, d[f"grid_sweep{sweep}_{k}"]), (sweep, k)
    assert np.all(st["eta"][:, 3] == 0)          # inactive column untouched


def _check_em(model, d, tag, n_iter, tol):
    for it in (1, n_iter):
        for c in model.shapes:
            for name in ("var_gamma", "var_mu", "eta", "q", "eta_diff", "zeta"):
                ref = d[f"{tag}_it{it}_{c}_{name}"]
                got = model.snapshots[it][c][name]
                assert got.shape == ref.shape
                assert relmax(got, ref) <= tol, (tag, it, c, name, relmax(got, ref))


def _run_oracle_em(model, theta, n_iter):
    model.snapshots = {}
    model.initialize(dict(theta))
    hist = {k: [] for k in ("elbo", "pi", "tau_beta", "sigma_epsilon", "sigma_g", "max_eta_diff", "mse")}
    for it in range(1, n_iter + 1):
        model.e_step()
        model.m_step()
        hist["elbo"].append(float(model.elbo()))
        hist["pi"].append(np.array(model.pi, dtype=np.float64))
        hist["tau_beta"].append(np.array(model.tau_beta, dtype=np.float64))
        hist["sigma_epsilon"].append(float(model.sigma_epsilon))
        hist["sigma_g"].append(float(model._sigma_g))
        hist["max_eta_diff"].append(float(max(np.max(np.abs(x)) for x in model.eta_diff.values())))
        hist["mse"].append(float(model.mse()))
        model.snapshots[it] = {c: {n: np.array(getattr(model, n)[c]) for n in
                                   ("var_gamma", "var_mu", "eta", "q", "eta_diff", "zeta")} for c in model.shapes}
    return hist


def _inputs(chroms):
    ld = {c: (v["ld_data"], v["ld_indptr"], v["ld_left_bound"]) for c, v in chroms.items()}
    return ld, {c: v["std_beta"] for c, v in chroms.items()}, {c: v["n_per_snp"] for c, v in chroms.items()}


@pytest.mark.parametrize("fname,prec,dq,tag,theta,fix", [
    ("viprs_f32_f32.npz", "float32", False, "em", {"pi": 0.05, "sigma_epsilon": 0.7}, None),
    ("viprs_f32_f32.npz", "float32", False, "fixeps", {"pi": 0.05}, {"sigma_epsilon": 0.75}),
    ("viprs_f64_f64.npz", "float64", False, "em", {"pi": 0.05, "sigma_epsilon": 0.7}, None),
    ("viprs_f32_i8.npz", "float32", True, "em", {"pi": 0.05, "sigma_epsilon": 0.7}, None),
])
def test_oracle_viprs_em_matches_reference_python(oracle_built, fname, prec, dq, tag, theta, fix):
    """OracleVIPRS (numpy restatement + C port) reproduces the reference's VIPRS class bit for bit."""
    d, chroms = load_golden(fname)
    ld, beta, n = _inputs(chroms)
    m = cpu.OracleVIPRS(ld, beta, n, fix_params=fix, float_precision=prec, dequantize_on_the_fly=dq, kind="port")
    hist = _run_oracle_em(m, theta, 5)
    _check_em(m, d, tag, 5, 0.0)
    for k, v in hist.items():
        ref = d[f"{tag}_hist_{k}"]
        assert np.allclose(np.array(v, dtype=np.float64), ref, rtol=1e-13, atol=0), (k, v, ref)


def test_oracle_viprsmix_em_matches_reference_python(oracle_built):
    d, chroms = load_golden("viprsmix_f32_i16.npz")
    ld, beta, n = _inputs(chroms)
    m = cpu.OracleVIPRSMix(ld, beta, n, K=4, float_precision="float32", dequantize_on_the_fly=True, kind="port")
    assert np.array_equal(m.d, d["mix_d"])
    hist = _run_oracle_em(m, {"pis": d["mix_pis"].copy(), "sigma_epsilon": 0.7}, 5)
    _check_em(m, d, "em", 5, 0.0)
    for k, v in hist.items():
        assert np.allclose(np.array(v, dtype=np.float64), d[f"em_hist_{k}"], rtol=1e-13, atol=0), k
