"""
GPU tests of the device-resident EM path (viprs_b200/model.py: VIPRS, VIPRSMix, VIPRSGrid; the prepare / sums
kernels) against the golden vectors the reference's own VIPRS / VIPRSMix classes produced
(tests/golden/make_golden.py) and against the oracle.  Tolerances: 1e-4 relative for float32 state, 1e-10 for
float64 (BASELINE.json), max-norm relative for arrays.
"""
import numpy as np
import pytest

from conftest import load_golden, relmax
from tests_util import make_block_ld

pytestmark = pytest.mark.gpu

N_ITER = 5


@pytest.fixture(scope="module")
def vb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import viprs_b200
    return viprs_b200


def _run_em(model, theta, n_iter):
    """The body of VIPRS.fit's loop for a fixed iteration count -- what make_golden.run_model does with the reference."""
    model.initialize(dict(theta))
    hist = {k: [] for k in ("elbo", "pi", "tau_beta", "sigma_epsilon", "sigma_g", "max_eta_diff", "mse")}
    snaps = {}
    for it in range(n_iter):
        model.e_step()
        model.m_step()
        hist["elbo"].append(model.elbo())
        hist["pi"].append(np.array(model.pi, dtype=np.float64))
        hist["tau_beta"].append(np.array(model.tau_beta, dtype=np.float64))
        hist["sigma_epsilon"].append(model.sigma_epsilon)
        hist["sigma_g"].append(model._sigma_g)
        hist["max_eta_diff"].append(model.max_eta_diff())
        hist["mse"].append(model.mse())
        if it in (0, n_iter - 1):
            snaps[it + 1] = {name: {c: v.detach().cpu().numpy().copy() for c, v in getattr(model, name).items()}
                             for name in ("var_gamma", "var_mu", "eta", "q", "eta_diff", "zeta")}
    return {k: np.array(v) for k, v in hist.items()}, snaps


def _check_against_golden(d, hist, snaps, tol, tag="em"):
    for k in ("elbo", "pi", "tau_beta", "sigma_epsilon", "sigma_g", "mse"):
        ref = d[f"{tag}_hist_{k}"]
        assert np.allclose(hist[k], ref, rtol=tol, atol=0), (k, hist[k], ref)
    assert np.allclose(hist["max_eta_diff"], d[f"{tag}_hist_max_eta_diff"], rtol=max(tol, 1e-6) * 10)
    for it, snap in snaps.items():
        for name, per_chrom in snap.items():
            for c, a in per_chrom.items():
                ref = d[f"{tag}_it{it}_{c}_{name}"]
                if name == "eta_diff":
                    # a difference of two nearly equal etas: near convergence its float32 rounding noise is of the
                    # order of eps * |eta| whatever the evaluation order, so it is judged on the scale of eta
                    scale = np.max(np.abs(d[f"{tag}_it{it}_{c}_eta"]))
                    err = np.max(np.abs(a.astype(np.float64) - ref)) / scale
                    assert err <= tol, (it, name, c, err)
                    continue
                assert relmax(a, ref) <= tol, (it, name, c, relmax(a, ref))


@pytest.mark.parametrize("name,prec,tol", [("viprs_f32_f32.npz", "float32", 1e-4), ("viprs_f32_i8.npz", "float32", 1e-4),
                                           ("viprs_f64_f64.npz", "float64", 1e-10)])
def test_viprs_em_matches_reference_golden(vb, name, prec, tol):
    from viprs_b200.model import VIPRS
    d, ch = load_golden(name)
    m = VIPRS(data={c: ch[c] for c in sorted(ch)}, float_precision=prec)
    hist, snaps = _run_em(m, {"pi": 0.05, "sigma_epsilon": 0.7}, N_ITER)
    _check_against_golden(d, hist, snaps, tol)


def test_viprs_fixed_sigma_epsilon_and_fit_match_golden(vb):
    from viprs_b200.model import VIPRS
    d, ch = load_golden("viprs_f32_f32.npz")
    data = {c: ch[c] for c in sorted(ch)}
    m = VIPRS(data=data, fix_params={"sigma_epsilon": 0.75}, float_precision="float32")
    hist, snaps = _run_em(m, {"pi": 0.05}, N_ITER)
    _check_against_golden(d, hist, snaps, 1e-4, tag="fixeps")
    # fit(): the reference's own convergence logic on the same inputs (8 iterations max)
    m2 = VIPRS(data=data, float_precision="float32")
    m2.fit(max_iter=8, theta_0={"pi": 0.05, "sigma_epsilon": 0.7})
    ref = d["fit_elbo_history"]
    assert len(m2.history["ELBO"]) == len(ref)
    assert np.allclose(m2.history["ELBO"], ref, rtol=1e-4)
    for c in data:
        assert relmax(m2.pip[c], d[f"fit_{c}_pip"]) <= 1e-4
        assert relmax(m2.post_mean_beta[c], d[f"fit_{c}_post_mean_beta"]) <= 1e-4
        assert relmax(m2.post_var_beta[c], d[f"fit_{c}_post_var_beta"]) <= 1e-4
    assert m2.optim_result.stop_iteration


def test_viprsmix_em_matches_reference_golden(vb):
    from viprs_b200.model import VIPRSMix
    d, ch = load_golden("viprsmix_f32_i16.npz")
    m = VIPRSMix(data={c: ch[c] for c in sorted(ch)}, K=4, float_precision="float32")
    assert np.allclose(m.d, d["mix_d"])
    hist, snaps = _run_em(m, {"pis": d["mix_pis"].copy(), "sigma_epsilon": 0.7}, N_ITER)
    _check_against_golden(d, hist, snaps, 1e-4)


def test_prepare_and_sums_kernels_match_numpy(vb, oracle_built):
    """viprs_b200_prepare_* / viprs_b200_sums_* against the numpy restatements, float32 and float64, all layouts."""
    import torch
    from oracle import cpu as ocpu
    from viprs_b200 import _lib
    from viprs_b200.ld import _stream_ptr
    L = _lib.lib()
    rng = np.random.default_rng(3)
    M = 5000
    seg = np.array([0, 1200, 1200, 3100, M], dtype=np.int32)          # one empty segment on purpose
    n = np.floor(rng.uniform(4e4, 6e4, M))
    for T, tt, rtol in ((np.float32, torch.float32, 2e-6), (np.float64, torch.float64, 1e-13)):
        for layout, ncol in ((0, 1), (0, 7), (1, 4)):
            theta = np.stack([rng.uniform(0.5, 0.9, ncol), rng.uniform(50, 500, ncol), rng.uniform(0.005, 0.1, ncol),
                              np.full(ncol, 0.01)], axis=1)
            if layout == 1:
                theta[:, 0] = theta[0, 0]
            shape = (M, ncol)
            g = rng.uniform(0, 1, shape).astype(T)
            g[::97] = 0.0
            g[5::101] = 1.0 if layout == 0 else g[5::101]
            if layout == 1:
                g = (g / ncol).astype(T)
            mu = (rng.standard_normal(shape) * 1e-2).astype(T)
            vshape = shape if layout == 0 else (M,)
            eta, q, diff = [(rng.standard_normal(vshape) * 1e-2).astype(T) for _ in range(3)]
            beta = (rng.standard_normal(M) * 1e-2).astype(T)
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.T if (layout == 0 and a.ndim == 2) else a)).cuda()
            d_g, d_mu, d_eta, d_q, d_diff, d_beta = dev(g), dev(mu), dev(eta), dev(q), dev(diff), dev(beta)
            d_n, d_th, d_seg = torch.from_numpy(n).cuda(), torch.from_numpy(theta).cuda(), torch.from_numpy(seg).cuda()
            nseg = len(seg) - 1
            ws = torch.zeros(int(L.viprs_b200_sums_workspace_bytes(M, ncol, nseg)), dtype=torch.uint8, device="cuda")
            out = torch.zeros((nseg, ncol, _lib.NSUMS), dtype=torch.float64, device="cuda")
            fn = L.viprs_b200_sums_f32 if T == np.float32 else L.viprs_b200_sums_f64
            for rep in range(2):                      # twice: the workspace counters reset themselves
                rc = fn(M, ncol, layout, nseg, d_seg.data_ptr(), d_g.data_ptr(), d_mu.data_ptr(), d_eta.data_ptr(),
                        d_q.data_ptr(), d_diff.data_ptr(), d_beta.data_ptr(), d_n.data_ptr(), d_th.data_ptr(), None, 2.0,
                        ws.data_ptr(), ws.numel(), out.data_ptr(), _stream_ptr())
                assert rc == 0
                got = out.cpu().numpy()
                for s in range(nseg):
                    a, b = seg[s], seg[s + 1]
                    ref = ocpu.sums_numpy(g[a:b], mu[a:b], eta[a:b], q[a:b], diff[a:b], beta[a:b], n[a:b], theta,
                                          q_scale=2.0, mixture=(layout == 1)) if b > a else np.zeros((ncol, 16))
                    assert np.allclose(got[s], ref, rtol=rtol, atol=1e-12), (T, layout, ncol, s)
            # prepare
            fnp = L.viprs_b200_prepare_f32 if T == np.float32 else L.viprs_b200_prepare_f64
            o = [torch.zeros_like(d_g) for _ in range(3)]
            lnp = torch.zeros(M, dtype=tt, device="cuda") if layout == 1 else None
            for half in (0, 1):
                rc = fnp(M, ncol, layout, half, d_n.data_ptr(), d_th.data_ptr(), o[0].data_ptr(), o[1].data_ptr(),
                         o[2].data_ptr(), lnp.data_ptr() if lnp is not None else None, _stream_ptr())
                assert rc == 0
                vt = n[:, None] * (1 + theta[:, 3]) / theta[:, 0] + theta[:, 1]
                ul = np.log(theta[:, 2]) - np.log(1 - theta[:, 2]) + .5 * (np.log(theta[:, 1]) - np.log(vt))
                tt_ref = .5 * vt if half else np.sqrt(.5 * vt)
                mm = n[:, None] / (vt * theta[:, 0])
                back = lambda t: t.cpu().numpy().T if layout == 0 else t.cpu().numpy()
                for a, b in ((back(o[0]), ul), (back(o[1]), tt_ref), (back(o[2]), mm)):
                    assert np.allclose(a.reshape(M, ncol), b.astype(T), rtol=4e-7 if T == np.float32 else 1e-15)
                if lnp is not None:
                    assert np.allclose(lnp.cpu().numpy(), np.log(1 - theta[:, 2].sum()), rtol=1e-6)


def _oracle_grid_column(ocpu):
    """The reference's VIPRS host arithmetic (numpy restatement) driving the C++ e_step_grid on ONE column: the
    semantics a column of cpp_e_step_grid has (no skip branch, half_var_tau, mu_mult * (beta - q); e_step.hpp:599-634)."""
    class OracleGridColumn(ocpu.OracleVIPRS):
        def e_step(self):
            for c in self.shapes:
                tau_beta, pi = self.tau_beta, self.pi
                self.var_tau[c] = (self.n_per_snp[c] * (1. + self.lambda_min) / self.sigma_epsilon) + tau_beta
                np.log(self.var_tau[c], out=self._log_var_tau[c])
                fp = self.float_precision
                col = lambda a: np.asfortranarray(np.asarray(a, dtype=fp).reshape(-1, 1))
                mu_mult = col(self.n_per_snp[c] / (self.var_tau[c] * self.sigma_epsilon))
                u_logs = col(np.log(pi) - np.log(1. - pi) + .5 * (np.log(tau_beta) - self._log_var_tau[c]))
                st = {k: col(getattr(self, k)[c]) for k in ("var_gamma", "var_mu", "eta", "q", "eta_diff")}
                ocpu.e_step_grid(self.ld_left_bound[c], self.ld_indptr[c], self.ld_data[c], self.std_beta[c],
                                 st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"], u_logs,
                                 col(0.5 * self.var_tau[c]), mu_mult, self.dequantize_scale, np.zeros(1, np.int32),
                                 self.threads, self.low_memory, kind=self.kind)
                for k, v in st.items():
                    getattr(self, k)[c][:] = v[:, 0]
            self.zeta = self.compute_zeta()
    return OracleGridColumn


def test_viprsgrid_batched_matches_independent_reference_fits(vb, oracle_built):
    """VIPRSGrid.fit(pathwise=False): every grid column must land where an independent per-column fit with the
    reference's e_step_grid sweep and VIPRS's M-step lands after the same number of EM iterations."""
    from oracle import cpu as ocpu
    from viprs_b200.model import VIPRSGrid
    d, ch = load_golden("viprs_f32_i8.npz")
    data = {c: ch[c] for c in sorted(ch)}
    grid = [{"pi": p, "sigma_epsilon": s} for s in (0.6, 0.85) for p in (0.005, 0.02, 0.1)]
    m = VIPRSGrid(data=data, grid=grid, float_precision="float32")
    n_it = 6
    m.fit(pathwise=False, max_iter=n_it, min_iter=100)
    assert len(m.history["ELBO"]) == n_it + 1 and m.optim_result.nit == n_it * len(grid)
    keys = sorted(ch)
    Oracle = _oracle_grid_column(ocpu)
    for g, rec in enumerate(grid):
        o = Oracle({c: (ch[c]["ld_data"], ch[c]["ld_indptr"], ch[c]["ld_left_bound"]) for c in keys},
                   {c: ch[c]["std_beta"] for c in keys}, {c: ch[c]["n_per_snp"] for c in keys},
                   fix_params=dict(rec), float_precision="float32", dequantize_on_the_fly=True)
        o.run(n_it, {})
        assert np.isclose(m.history["ELBO"][-1][g], o.history["ELBO"][-1], rtol=1e-4), (g, rec)
        assert np.isclose(m.tau_beta[g], float(o.tau_beta), rtol=1e-4)
        for c in keys:
            assert relmax(m.var_gamma[c][:, g].cpu().numpy(), o.var_gamma[c]) <= 1e-4, (g, c)
            assert relmax(m.eta[c][:, g].cpu().numpy(), o.eta[c]) <= 1e-4, (g, c)
            assert relmax(m.q[c][:, g].cpu().numpy(), o.q[c]) <= 1e-4, (g, c)
    assert m.pip[keys[0]].shape == (len(ch[keys[0]]["std_beta"]), len(grid))
    # convergence handling: with the default stopping rules columns stop independently and drop out of the sweep
    m2 = VIPRSGrid(data=data, grid=grid, float_precision="float32")
    m2.fit(pathwise=False, max_iter=400)
    assert all(o.stop_iteration for o in m2.optim_results)
    nits = [o.nit for o in m2.optim_results]
    assert len(set(nits)) > 1 and max(nits) < 400
    assert m2.converged_models.all()


def test_viprsgrid_pathwise_matches_serial_warm_started_reference(vb, oracle_built):
    """fit(pathwise=True) == the reference's serial loop: fix the next grid point, continue from the current state."""
    from viprs_b200.model import VIPRS, VIPRSGrid
    d, ch = load_golden("viprs_f32_i8.npz")
    data = {c: ch[c] for c in sorted(ch)}
    grid = [{"pi": 0.01, "sigma_epsilon": 0.8}, {"pi": 0.03, "sigma_epsilon": 0.8}, {"pi": 0.1, "sigma_epsilon": 0.8}]
    m = VIPRSGrid(data=data, grid=grid, float_precision="float32")
    m.fit(pathwise=True, max_iter=15)
    # the same thing spelled out with the single-model class
    s = VIPRS(data=data, float_precision="float32", fix_params=dict(grid[0]))
    elbos = []
    for i, rec in enumerate(grid):
        if i > 0:
            s.set_fixed_params(rec)
        s.fit(max_iter=15, continued=i > 0)
        elbos.append(s.history["ELBO"][-1])
        s.optim_result.reset()
        c = sorted(ch)[0]
        assert relmax(m.var_gamma[c][:, i].cpu().numpy(), s.var_gamma[c].cpu().numpy()) <= 1e-6
    assert np.allclose([r["ELBO"] for r in m.validation_result], elbos, rtol=1e-9)
    assert m.pip[sorted(ch)[0]].shape[1] == 3


def test_model_on_bigger_blocks_against_oracle(vb, oracle_built):
    """Two chromosomes with blocks up to 1500 SNPs, int16 LD: 4 EM iterations against the numpy restatement driving the
    compiled reference."""
    from oracle import cpu as ocpu
    from viprs_b200.model import VIPRS
    rng = np.random.default_rng(21)
    data = {}
    for c, sizes in ((1, (1500, 300, 77)), (2, (900, 1, 640))):
        P = make_block_ld(rng, sizes, np.int16, np.float32)
        data[c] = dict(ld_data=P["data"], ld_indptr=P["indptr"], ld_left_bound=P["lb"], std_beta=P["beta"],
                       n_per_snp=np.floor(rng.uniform(4e4, 6e4, P["M"])))
    theta = {"pi": 0.02, "sigma_epsilon": 0.8}
    m = VIPRS(data=data, float_precision="float32")
    hist, _ = _run_em(m, theta, 4)
    o = ocpu.OracleVIPRS({c: (v["ld_data"], v["ld_indptr"], v["ld_left_bound"]) for c, v in data.items()},
                         {c: v["std_beta"] for c, v in data.items()}, {c: v["n_per_snp"] for c, v in data.items()},
                         float_precision="float32", dequantize_on_the_fly=True)
    o.run(4, dict(theta))
    assert np.allclose(hist["elbo"], o.history["ELBO"], rtol=1e-4)
    assert np.allclose(hist["sigma_epsilon"], o.history["sigma_epsilon"], rtol=1e-4)
    for c in data:
        assert relmax(m.eta[c].cpu().numpy(), o.eta[c]) <= 1e-4
        assert relmax(m.q[c].cpu().numpy(), o.q[c]) <= 1e-4


@pytest.mark.parametrize("pathwise", [False, True])
def test_viprsgrid_lambda_min_column_matches_oracle(vb, oracle_built, pathwise):
    """A grid with a lambda_min column (the reference's set_fixed_params sets self.lambda_min for every grid point,
    VIPRS.py:855-870 / VIPRSGrid.py:190-196): every column against an independent fit of the numpy restatement with that
    lambda_min (advisor finding, round 1: model 0 of a pathwise fit used the constructor's lambda_min)."""
    from oracle import cpu as ocpu
    from viprs_b200.model import VIPRSGrid
    d, ch = load_golden("viprs_f32_i8.npz")
    keys = sorted(ch)
    data = {c: ch[c] for c in keys}
    grid = [{"pi": 0.02, "sigma_epsilon": 0.8, "lambda_min": lm} for lm in (0.3, 0.0, 0.1)]
    m = VIPRSGrid(data=data, grid=grid, float_precision="float32")
    n_it = 5
    if pathwise:
        m.fit(pathwise=True, max_iter=n_it, min_iter=100)
    else:
        m.fit(pathwise=False, max_iter=n_it, min_iter=100)
    Oracle = _oracle_grid_column(ocpu) if not pathwise else ocpu.OracleVIPRS
    o = None
    for g, rec in enumerate(grid):
        fix = {k: v for k, v in rec.items() if k != "lambda_min"}
        if not pathwise or o is None:
            o = Oracle({c: (ch[c]["ld_data"], ch[c]["ld_indptr"], ch[c]["ld_left_bound"]) for c in keys},
                       {c: ch[c]["std_beta"] for c in keys}, {c: ch[c]["n_per_snp"] for c in keys},
                       fix_params=fix, lambda_min=rec["lambda_min"], float_precision="float32", dequantize_on_the_fly=True)
            o.run(n_it, {})
        else:
            # the reference's pathwise loop: fix the next grid point and continue from the current state
            o.lambda_min = rec["lambda_min"]
            o.fix_params.update(fix)
            o.sigma_epsilon, o.pi = fix["sigma_epsilon"], fix["pi"]
            for _ in range(n_it):
                o.e_step(); o.m_step()
        for c in keys:
            assert relmax(m.var_gamma[c][:, g].cpu().numpy(), o.var_gamma[c]) <= 2e-4, (g, c, pathwise)
            assert relmax(m.eta[c][:, g].cpu().numpy(), o.eta[c]) <= 2e-4, (g, c, pathwise)
