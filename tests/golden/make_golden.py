"""
Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, shz9/viprs v0.1.4) in the build container:

  * the reference's Cython layer viprs/model/vi/e_step_cpp.pyx is cythonized and compiled from a
    scratch copy under /tmp (the reference tree is read-only; nothing is copied into this repo),
  * `magenpy` (absent here, no network) is replaced by a stub that only provides the names the
    reference imports; the LD matrices / summary statistics come from a tiny stub GWADataLoader that
    hands the reference the same raw arrays magenpy would (VIPRS.py:153-172, BayesPRSModel.py:133-136),
  * the reference's own `VIPRS` / `VIPRSMix` classes run `initialize(); [e_step(); m_step(); elbo()] * n`
    (the body of `VIPRS.fit`, VIPRS.py:979-994) and `fit()`; `cpp_e_step_grid` is called directly
    (nothing in the reference's Python reaches it).

Each .npz holds the inputs and the reference's outputs, so the tests need neither /root/reference
nor this script at run time.      Usage:  python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SCRATCH = "/tmp/viprs_ref_build"


def build_reference_extension():
    if not os.path.exists(os.path.join(SCRATCH, "viprs", "model", "vi")):
        shutil.copytree(os.path.join(REF, "viprs"), os.path.join(SCRATCH, "viprs"))
    vi = os.path.join(SCRATCH, "viprs", "model", "vi")
    if any(f.startswith("e_step_cpp") and f.endswith(".so") for f in os.listdir(vi)):
        return
    setup = os.path.join(SCRATCH, "setup_min.py")
    with open(setup, "w") as fh:
        fh.write(
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy as np\n"
            "ext = Extension('viprs.model.vi.e_step_cpp', ['viprs/model/vi/e_step_cpp.pyx'], language='c++',\n"
            "                include_dirs=[np.get_include(), 'viprs/model/vi'],\n"
            "                extra_compile_args=['-O3', '-std=c++17', '-fopenmp'], extra_link_args=['-fopenmp'])\n"
            "setup(name='viprs_ref_ext', ext_modules=cythonize([ext], language_level=3))\n")
    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++", LDSHARED="/usr/bin/g++ -shared")
    subprocess.run([sys.executable, setup, "build_ext", "--inplace"], cwd=SCRATCH, env=env, check=True,
                   stdout=subprocess.DEVNULL)


def stub_magenpy():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    mg = mod("magenpy")
    mg.GWADataLoader = type("GWADataLoader", (), {})
    mod("magenpy.utils")
    cu = mod("magenpy.utils.compute_utils")
    cu.is_numeric = lambda x: isinstance(x, (int, float, np.number, np.ndarray))
    mod("magenpy.stats")
    mod("magenpy.stats.h2")
    ldsc = mod("magenpy.stats.h2.ldsc")

    def simple_ldsc(gdl):
        raise RuntimeError("ldsc is not available in the stub")
    ldsc.simple_ldsc = simple_ldsc


class _LDLop:
    def __init__(self, data, indptr, lb):
        self.ld_data, self.ld_indptr, self.leftmost_idx = data, indptr, lb


class _LDMat:
    def __init__(self, data, indptr, lb):
        self._d, self._ip, self._lb = data, indptr, lb
        self.stored_dtype = data.dtype

    def load(self, return_symmetric=False, dtype=None):
        assert not return_symmetric
        d = self._d if np.dtype(dtype) == self._d.dtype else self._d.astype(dtype)
        return _LDLop(d, self._ip, self._lb)


class _SS:
    def __init__(self, beta, n):
        self.n_per_snp, self._b = n, beta

    def get_snp_pseudo_corr(self):
        return self._b


def StubGDL(chroms):
    """An instance of the stubbed magenpy.GWADataLoader (BayesPRSModel.py:49 asserts the type) carrying the
    few attributes the reference touches on this path."""
    cls = type("StubGDL", (sys.modules["magenpy"].GWADataLoader, _StubGDL), {})
    return cls(chroms)


class _StubGDL:
    def __init__(self, chroms):
        self._c = chroms
        self.genotype = None
        self.ld = True          # only tested against None (BayesPRSModel.py:51)
        self.shapes = {c: len(v["std_beta"]) for c, v in chroms.items()}
        self.sumstats_table = {c: _SS(v["std_beta"], v["n_per_snp"]) for c, v in chroms.items()}
        self.m = sum(self.shapes.values())
        self.n = max(v["n_per_snp"].max() for v in chroms.values())

    def get_ld_matrices(self):
        return {c: _LDMat(v["ld_data"], v["ld_indptr"], v["ld_left_bound"]) for c, v in self._c.items()}


def synth_chrom(rng, block_sizes, ld_dtype, n=50000., k=24, alpha=0.5, h2=0.4, p_causal=0.05, M_total=None):
    """Block-diagonal PD LD, upper-triangular CSR-without-column-indices (numpy only; seeded)."""
    data, lens, lbs, betas = [], [], [], []
    row0 = 0
    M_total = M_total or sum(block_sizes)
    for B in block_sizes:
        Z = rng.standard_normal((B, k))
        C = Z @ Z.T / k
        dinv = 1. / np.sqrt(np.diag(C))
        R = alpha * C * dinv[:, None] * dinv[None, :]
        np.fill_diagonal(R, 1.)
        bt = rng.standard_normal(B) * np.sqrt(h2 / (p_causal * M_total)) * (rng.random(B) < p_causal)
        Lc = np.linalg.cholesky(R)
        betas.append(R @ bt + Lc @ rng.standard_normal(B) / np.sqrt(n))
        if ld_dtype == "int8":
            Rq = np.rint(R * 127.).astype(np.int8)
        elif ld_dtype == "int16":
            Rq = np.rint(R * 32767.).astype(np.int16)
        else:
            Rq = R.astype(ld_dtype)
        iu = np.triu_indices(B, 1)
        data.append(Rq[iu])
        lens.append(np.arange(B - 1, -1, -1, dtype=np.int64))
        lbs.append(np.arange(row0 + 1, row0 + B + 1, dtype=np.int32))
        row0 += B
    indptr = np.zeros(row0 + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.concatenate(lens))
    nps = np.full(row0, n) * (1. + 0.1 * rng.random(row0))      # non-constant n_per_snp on purpose
    return {"ld_data": np.concatenate(data), "ld_indptr": indptr, "ld_left_bound": np.concatenate(lbs),
            "std_beta": np.concatenate(betas), "n_per_snp": np.floor(nps)}


def flat(chroms, prefix="in"):
    out = {}
    for c, v in chroms.items():
        for k_, a in v.items():
            out[f"{prefix}_{c}_{k_}"] = a
    return out


def run_model(model, theta_0, n_iter, out, tag):
    model.initialize(dict(theta_0))
    hist = {k_: [] for k_ in ("elbo", "pi", "tau_beta", "sigma_epsilon", "sigma_g", "max_eta_diff", "mse")}
    for it in range(n_iter):
        model.e_step()
        model.m_step()
        hist["elbo"].append(np.float64(model.elbo()))
        hist["pi"].append(np.array(model.pi, dtype=np.float64))
        hist["tau_beta"].append(np.array(model.tau_beta, dtype=np.float64))
        hist["sigma_epsilon"].append(np.float64(model.sigma_epsilon))
        hist["sigma_g"].append(np.float64(model._sigma_g))
        hist["max_eta_diff"].append(np.float64(max(np.max(np.abs(d)) for d in model.eta_diff.values())))
        hist["mse"].append(np.float64(model.mse()))
        if it in (0, n_iter - 1):
            for c in model.shapes:
                for name in ("var_gamma", "var_mu", "eta", "q", "eta_diff", "zeta"):
                    out[f"{tag}_it{it + 1}_{c}_{name}"] = np.array(getattr(model, name)[c])
    for k_, v in hist.items():
        out[f"{tag}_hist_{k_}"] = np.array(v)


def main():
    build_reference_extension()
    stub_magenpy()
    sys.path.insert(0, SCRATCH)
    from viprs.model.VIPRS import VIPRS
    from viprs.model.VIPRSMix import VIPRSMix
    from viprs.model.vi.e_step_cpp import cpp_e_step_grid, cpp_e_step, cpp_e_step_mixture

    N_ITER = 5
    # ---- case 1: VIPRS float32 state, float32 LD, two chromosomes, all hyper-parameters updated ----
    rng = np.random.Generator(np.random.Philox(key=7209))
    chroms = {21: synth_chrom(rng, [70, 45, 96], "float32", M_total=330),
              22: synth_chrom(rng, [64, 55], "float32", M_total=330)}
    out = flat(chroms)
    theta = {"pi": 0.05, "sigma_epsilon": 0.7}
    m = VIPRS(StubGDL(chroms), float_precision="float32", low_memory=True, threads=1)
    run_model(m, theta, N_ITER, out, "em")
    out["theta_pi"], out["theta_sigma_epsilon"] = 0.05, 0.7
    # reference fit() on the same inputs (runs its own convergence logic)
    m2 = VIPRS(StubGDL(chroms), float_precision="float32", low_memory=True, threads=1)
    m2.fit(max_iter=8, theta_0=dict(theta), disable_pbar=True)
    out["fit_elbo_history"] = np.array(m2.history["ELBO"], dtype=np.float64)
    for c in m2.shapes:
        out[f"fit_{c}_pip"] = m2.pip[c]
        out[f"fit_{c}_post_mean_beta"] = m2.post_mean_beta[c]
        out[f"fit_{c}_post_var_beta"] = m2.post_var_beta[c]
    # sigma_epsilon fixed (explicit likelihood term in the ELBO, VIPRS.py:549-558)
    m3 = VIPRS(StubGDL(chroms), fix_params={"sigma_epsilon": 0.75}, float_precision="float32", threads=1)
    run_model(m3, {"pi": 0.05}, N_ITER, out, "fixeps")
    np.savez_compressed(os.path.join(HERE, "viprs_f32_f32.npz"), **out)

    # ---- case 2: VIPRS float64 state, float64 LD ----
    rng = np.random.Generator(np.random.Philox(key=7210))
    chroms = {1: synth_chrom(rng, [80, 33, 120], "float64")}
    out = flat(chroms)
    m = VIPRS(StubGDL(chroms), float_precision="float64", low_memory=True, threads=1)
    run_model(m, {"pi": 0.05, "sigma_epsilon": 0.7}, N_ITER, out, "em")
    np.savez_compressed(os.path.join(HERE, "viprs_f64_f64.npz"), **out)

    # ---- case 3: VIPRS float32 state, int8 LD dequantised on the fly ----
    rng = np.random.Generator(np.random.Philox(key=7211))
    chroms = {7: synth_chrom(rng, [100, 61, 77], "int8")}
    out = flat(chroms)
    m = VIPRS(StubGDL(chroms), float_precision="float32", dequantize_on_the_fly=True, threads=1)
    assert abs(m.dequantize_scale - 1. / 127) < 1e-15
    run_model(m, {"pi": 0.05, "sigma_epsilon": 0.7}, N_ITER, out, "em")
    np.savez_compressed(os.path.join(HERE, "viprs_f32_i8.npz"), **out)

    # ---- case 4: VIPRSMix K=4, float32 state, int16 LD ----
    rng = np.random.Generator(np.random.Philox(key=7212))
    chroms = {3: synth_chrom(rng, [90, 50, 64], "int16")}
    out = flat(chroms)
    m = VIPRSMix(StubGDL(chroms), K=4, float_precision="float32", dequantize_on_the_fly=True, threads=1)
    out["mix_d"] = np.array(m.d)
    # "pis" is given explicitly: with only "pi" the reference draws a random Dirichlet split (VIPRSMix.py:84)
    out["mix_pis"] = 0.05 * np.array([0.4, 0.3, 0.2, 0.1])
    run_model(m, {"pis": out["mix_pis"].copy(), "sigma_epsilon": 0.7}, N_ITER, out, "em")
    np.savez_compressed(os.path.join(HERE, "viprsmix_f32_i16.npz"), **out)

    # ---- case 5: cpp_e_step_grid called directly, G=6 columns (one inactive), int8 LD ----
    rng = np.random.Generator(np.random.Philox(key=7213))
    ch = synth_chrom(rng, [72, 40, 88], "int8")
    M, G = len(ch["std_beta"]), 6
    pis = np.array([0.005, 0.02, 0.1, 0.005, 0.02, 0.1])
    sigs = np.array([0.6, 0.6, 0.6, 0.9, 0.9, 0.9])
    taus = pis * M / (1. - sigs)
    n = ch["n_per_snp"][:, None]
    var_tau = n / sigs[None, :] + taus[None, :]
    f32 = np.float32
    u_logs = np.asfortranarray((np.log(pis) - np.log(1 - pis) + .5 * (np.log(taus) - np.log(var_tau))).astype(f32))
    hvt = np.asfortranarray((0.5 * var_tau).astype(f32))
    mm = np.asfortranarray((n / (var_tau * sigs[None, :])).astype(f32))
    beta = ch["std_beta"].astype(f32)
    st = {k_: np.zeros((M, G), dtype=f32, order="F") for k_ in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.asfortranarray(np.tile(pis.astype(f32), (M, 1)))
    active = np.array([0, 1, 2, 4, 5], dtype=np.int32)
    out = flat({0: ch})
    out.update({"grid_u_logs": u_logs, "grid_half_var_tau": hvt, "grid_mu_mult": mm, "grid_active": active,
                "grid_pis": pis, "grid_sigma_epsilons": sigs, "grid_tau_betas": taus})
    for sweep in range(3):
        cpp_e_step_grid(ch["ld_left_bound"], ch["ld_indptr"], ch["ld_data"], beta, st["var_gamma"], st["var_mu"],
                        st["eta"], st["q"], st["eta_diff"], u_logs, hvt, mm, f32(1. / 127), active, 1, True)
        if sweep in (0, 2):
            for k_, a in st.items():
                out[f"grid_sweep{sweep + 1}_{k_}"] = a.copy(order="F")
    np.savez_compressed(os.path.join(HERE, "e_step_grid_f32_i8.npz"), **out)

    # ---- case 6: raw cpp_e_step / cpp_e_step_mixture single sweeps from a non-trivial state (f64, i16 LD) ----
    rng = np.random.Generator(np.random.Philox(key=7214))
    ch = synth_chrom(rng, [50, 81], "int16")
    M = len(ch["std_beta"])
    out = flat({0: ch})
    f64 = np.float64
    pi, se = 0.03, 0.8
    tau = pi * M / (1 - se)
    vt = ch["n_per_snp"] / se + tau
    args = {"u_logs": np.log(pi) - np.log(1 - pi) + .5 * (np.log(tau) - np.log(vt)),
            "sqrt_half_var_tau": np.sqrt(.5 * vt), "mu_mult": ch["n_per_snp"] / (vt * se)}
    st = {k_: np.zeros(M) for k_ in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.full(M, pi)
    for sweep in range(2):
        cpp_e_step(ch["ld_left_bound"], ch["ld_indptr"], ch["ld_data"], ch["std_beta"], st["var_gamma"], st["var_mu"],
                   st["eta"], st["q"], st["eta_diff"], args["u_logs"], args["sqrt_half_var_tau"], args["mu_mult"],
                   f64(1. / 32767), 1, True)
        for k_, a in st.items():
            out[f"raw_sweep{sweep + 1}_{k_}"] = a.copy()
    out.update({f"raw_{k_}": v for k_, v in args.items()})
    np.savez_compressed(os.path.join(HERE, "cpp_e_step_f64_i16.npz"), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
