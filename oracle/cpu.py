"""
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

ctypes bindings for the compiled checkers plus numpy restatements of the reference's
host-side EM arithmetic.  All citations are relative to /root/reference/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_TN = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_UN = {np.dtype(np.int8): "i8", np.dtype(np.int16): "i16",
       np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _path(kind):
    return os.path.join(_HERE, "_ref", "libviprs_ref.so" if kind == "reference" else "libviprs_port.so")


def have_ref():
    """True when the compiled *reference* (not just the port) is available."""
    return os.path.exists(_path("reference"))


def build(quiet=True):
    """Compile the checkers (`make -C oracle`).  Building the checker is not using it."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def load_lib(kind="auto"):
    """kind: 'reference' | 'port' | 'auto' (reference when built, else port)."""
    if kind == "auto":
        kind = "reference" if have_ref() else "port"
    if kind not in _LIBS:
        p = _path(kind)
        if not os.path.exists(p):
            build()
        _LIBS[kind] = ctypes.CDLL(p)
    return _LIBS[kind], kind


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(*arrs):
    for a in arrs:
        assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"], "oracle needs contiguous arrays"


def _scalar(T, v):
    return ctypes.c_float(v) if T == np.float32 else ctypes.c_double(v)


def e_step(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
           u_logs, sqrt_half_var_tau, mu_mult, dq_scale, threads=1, low_memory=True, kind="auto"):
    """Same argument order as cpp_e_step (viprs/model/vi/e_step_cpp.pyx:91-122); in-place."""
    lib, kind = load_lib(kind)
    T = var_mu.dtype.type
    name = f"e_step_{_TN[var_mu.dtype]}_{_UN[ld_data.dtype]}"
    lb = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
    ip = np.ascontiguousarray(ld_indptr, dtype=np.int64)
    _check(ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult)
    args = [ctypes.c_int(var_mu.shape[0]), _p(lb), _p(ip), _p(ld_data), _p(std_beta), _p(var_gamma),
            _p(var_mu), _p(eta), _p(q), _p(eta_diff), _p(u_logs), _p(sqrt_half_var_tau), _p(mu_mult),
            _scalar(T, dq_scale)]
    if kind == "reference":
        getattr(lib, "ref_" + name)(*args, ctypes.c_int(threads), ctypes.c_int(int(low_memory)))
    else:
        getattr(lib, "port_" + name)(*args, ctypes.c_int(int(low_memory)))


def e_step_mixture(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                   log_null_pi, u_logs, sqrt_half_var_tau, mu_mult, dq_scale, threads=1,
                   low_memory=True, kind="auto"):
    """cpp_e_step_mixture (e_step_cpp.pyx:125-159); (M,K) arrays C-order; in-place."""
    lib, kind = load_lib(kind)
    T = var_mu.dtype.type
    name = f"e_step_mixture_{_TN[var_mu.dtype]}_{_UN[ld_data.dtype]}"
    lb = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
    ip = np.ascontiguousarray(ld_indptr, dtype=np.int64)
    for a in (var_gamma, var_mu, u_logs, sqrt_half_var_tau, mu_mult):
        assert a.flags["C_CONTIGUOUS"] and a.ndim == 2
    args = [ctypes.c_int(var_mu.shape[0]), ctypes.c_int(var_mu.shape[1]), _p(lb), _p(ip), _p(ld_data),
            _p(std_beta), _p(var_gamma), _p(var_mu), _p(eta), _p(q), _p(eta_diff), _p(log_null_pi),
            _p(u_logs), _p(sqrt_half_var_tau), _p(mu_mult), _scalar(T, dq_scale)]
    if kind == "reference":
        getattr(lib, "ref_" + name)(*args, ctypes.c_int(threads), ctypes.c_int(int(low_memory)))
    else:
        getattr(lib, "port_" + name)(*args, ctypes.c_int(int(low_memory)))


def e_step_grid(ld_left_bound, ld_indptr, ld_data, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                u_logs, half_var_tau, mu_mult, dq_scale, active_model_idx, threads=1,
                low_memory=True, kind="auto"):
    """cpp_e_step_grid (e_step_cpp.pyx:161-195); (M,G) arrays Fortran-order; in-place."""
    lib, kind = load_lib(kind)
    T = var_mu.dtype.type
    name = f"e_step_grid_{_TN[var_mu.dtype]}_{_UN[ld_data.dtype]}"
    lb = np.ascontiguousarray(ld_left_bound, dtype=np.int32)
    ip = np.ascontiguousarray(ld_indptr, dtype=np.int64)
    act = np.ascontiguousarray(active_model_idx, dtype=np.int32)
    for a in (var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult):
        assert a.flags["F_CONTIGUOUS"] and a.ndim == 2
    args = [ctypes.c_int(var_mu.shape[0]), ctypes.c_int(act.shape[0]), _p(act), _p(lb), _p(ip),
            _p(ld_data), _p(std_beta), _p(var_gamma), _p(var_mu), _p(eta), _p(q), _p(eta_diff),
            _p(u_logs), _p(half_var_tau), _p(mu_mult), _scalar(T, dq_scale)]
    if kind == "reference":
        getattr(lib, "ref_" + name)(*args, ctypes.c_int(threads), ctypes.c_int(int(low_memory)))
    else:
        getattr(lib, "port_" + name)(*args, ctypes.c_int(int(low_memory)))


# ---------------------------------------------------------------------------------------------
# numpy restatement of the reference's host-side EM arithmetic
# ---------------------------------------------------------------------------------------------

def _dict_mean(d, axis=None):      # viprs/utils/compute_utils.py:43-49
    return np.mean(np.array([np.mean(v, axis=axis) for v in d.values()]), axis=axis)


def _dict_sum(d, axis=None):       # viprs/utils/compute_utils.py:52-62
    return np.sum(np.array([np.sum(v, axis=axis) for v in d.values()]), axis=axis)


def _dict_concat(d, axis=0):       # viprs/utils/compute_utils.py:22-31
    if len(d) == 1:
        return d[next(iter(d))]
    return np.concatenate([d[c] for c in sorted(d.keys())], axis=axis)


class OracleVIPRS:
    """
    Restatement of viprs/model/VIPRS.py for raw inputs (no magenpy):
      ld        : {chrom: (ld_data, ld_indptr, ld_left_bound)}   (VIPRS.py:167-172)
      std_beta  : {chrom: float array}                            (BayesPRSModel.py:135)
      n_per_snp : {chrom: array}                                  (BayesPRSModel.py:133)
    """

    def __init__(self, ld, std_beta, n_per_snp, fix_params=None, lambda_min=None,
                 float_precision="float32", low_memory=True, dequantize_on_the_fly=False,
                 threads=1, kind="auto"):
        self.float_precision = float_precision
        self.kind = kind
        self.ld_data = {c: v[0] for c, v in ld.items()}
        self.ld_indptr = {c: v[1] for c, v in ld.items()}
        self.ld_left_bound = {c: v[2] for c, v in ld.items()}
        self.shapes = {c: int(len(v[1]) - 1) for c, v in ld.items()}
        self.n_per_snp = {c: np.asarray(n) for c, n in n_per_snp.items()}
        self.std_beta = {c: np.asarray(b).astype(float_precision) for c, b in std_beta.items()}
        self._sample_size = np.max(np.array([np.max(v) for v in self.n_per_snp.values()]))  # BayesPRSModel.py:75
        self.lambda_min = 0. if lambda_min is None else lambda_min                         # VIPRS.py:177-181
        self.threads = threads
        self.fix_params = dict(fix_params or {})
        self.low_memory = low_memory
        first = self.ld_data[sorted(self.ld_data)[0]]
        if dequantize_on_the_fly and np.issubdtype(first.dtype, np.integer):                # VIPRS.py:203-207
            self.dequantize_scale = 1. / np.iinfo(first.dtype).max
        else:
            self.dequantize_scale = 1.
        self.history = {}

    @property
    def n_snps(self):
        return sum(self.shapes.values())

    @property
    def n(self):
        return self._sample_size

    # VIPRS.py:245-316 (the random branches are kept but tests always fix pi / sigma_epsilon)
    def initialize_theta(self, theta_0=None):
        if theta_0 is not None and self.fix_params is not None:
            theta_0 = dict(theta_0)
            theta_0.update(self.fix_params)
        elif self.fix_params is not None:
            theta_0 = self.fix_params
        elif theta_0 is None:
            theta_0 = {}
        if "pi" not in theta_0:
            self.pi = np.random.uniform(low=max(10. / self.n_snps, 1e-5), high=min(0.2, 1e4 / self.n_snps))
        else:
            self.pi = theta_0["pi"]
        if "sigma_epsilon" not in theta_0:
            if "tau_beta" not in theta_0:
                naive_h2g = np.random.uniform(low=.01, high=.1)
                self.sigma_epsilon = 1. - naive_h2g
                self.tau_beta = self.pi * self.n_snps / max(naive_h2g, 0.01)
            else:
                self.tau_beta = theta_0["tau_beta"]
                self.sigma_epsilon = np.clip(1. - (self.pi * self.n_snps / self.tau_beta), a_min=1e-4, a_max=1. - 1e-4)
        else:
            self.sigma_epsilon = theta_0["sigma_epsilon"]
            if "tau_beta" in theta_0:
                self.tau_beta = theta_0["tau_beta"]
            else:
                self.tau_beta = (self.pi * self.n_snps) / np.maximum(0.01, 1. - self.sigma_epsilon)
        ft = np.dtype(self.float_precision).type
        self.sigma_epsilon = ft(self.sigma_epsilon)
        self.pi = ft(self.pi)
        self.lambda_min = ft(self.lambda_min)
        self._sigma_g = ft(0.)

    # VIPRS.py:318-359 (param_0 may carry 'mu' / 'gamma', :339-351; q starts at 0 either way, :357)
    def initialize_variational_parameters(self, param_0=None):
        param_0 = param_0 or {}
        self.var_mu, self.var_tau, self.var_gamma = {}, {}, {}
        for c, shp in self._param_shapes().items():
            self.var_tau[c] = (self._n(c) / self.sigma_epsilon) + self.tau_beta
            if "mu" in param_0:
                self.var_mu[c] = np.array(param_0["mu"][c]).astype(self.float_precision)
            else:
                self.var_mu[c] = np.zeros(shp, dtype=self.float_precision)
            if "gamma" in param_0:
                self.var_gamma[c] = np.array(param_0["gamma"][c]).astype(self.float_precision)
            else:
                self.var_gamma[c] = self.pi * np.ones(shp, dtype=self.float_precision)
        self.eta = self.compute_eta()
        self.zeta = self.compute_zeta()
        self.eta_diff = {c: np.zeros_like(e, dtype=self.float_precision) for c, e in self.eta.items()}
        self.q = {c: np.zeros_like(e, dtype=self.float_precision) for c, e in self.eta.items()}
        self._log_var_tau = {c: np.log(self.var_tau[c]) for c in self.var_tau}

    def _param_shapes(self):
        return self.shapes

    def _n(self, c):
        return self.n_per_snp[c]

    def initialize(self, theta_0=None, param_0=None):
        self.initialize_theta(theta_0)
        self.initialize_variational_parameters(param_0) if param_0 else self.initialize_variational_parameters()
        self.history = {"ELBO": [], "sigma_epsilon": [], "tau_beta": [], "pi": [], "sigma_g": [],
                        "max_eta_diff": [], "mse": []}

    # VIPRS.py:381-424
    def e_step(self):
        for c in self.shapes:
            tau_beta, pi = self.tau_beta, self.pi
            self.var_tau[c] = (self.n_per_snp[c] * (1. + self.lambda_min) / self.sigma_epsilon) + tau_beta
            np.log(self.var_tau[c], out=self._log_var_tau[c])
            mu_mult = (self.n_per_snp[c] / (self.var_tau[c] * self.sigma_epsilon)).astype(self.float_precision)
            u_logs = (np.log(pi) - np.log(1. - pi) + .5 * (np.log(tau_beta) - self._log_var_tau[c])
                      ).astype(self.float_precision)
            e_step(self.ld_left_bound[c], self.ld_indptr[c], self.ld_data[c], self.std_beta[c],
                   self.var_gamma[c], self.var_mu[c], self.eta[c], self.q[c], self.eta_diff[c],
                   u_logs, np.sqrt(0.5 * self.var_tau[c]).astype(self.float_precision), mu_mult,
                   self.dequantize_scale, self.threads, self.low_memory, kind=self.kind)
        self.zeta = self.compute_zeta()

    # VIPRS.py:426-484
    def m_step(self):
        if "pi" not in self.fix_params:
            self.pi = _dict_mean(self.var_gamma, axis=0)
        if "tau_beta" not in self.fix_params:
            self.tau_beta = (self.pi * self.n_snps / _dict_sum(self.zeta, axis=0))
        self._sigma_g = np.sum([
            np.sum((1. + self.lambda_min) * self.zeta[c] + np.multiply(self.q[c], self.eta[c]), axis=0)
            for c in self.shapes.keys()], axis=0)
        if "sigma_epsilon" not in self.fix_params:
            sig_eps = 0.
            for c in self.shapes:
                sig_eps -= 2. * self.std_beta[c].dot(self.eta[c])
            self.sigma_epsilon = 1. + sig_eps + self._sigma_g

    def compute_pip(self):                               # VIPRS.py:875-880
        return self.var_gamma.copy()

    def compute_eta(self):                               # VIPRS.py:882-886
        return {c: v * self.var_mu[c] for c, v in self.var_gamma.items()}

    def compute_zeta(self):                              # VIPRS.py:888-897
        return {c: np.multiply(v, self.var_mu[c].astype(np.float64) ** 2 + 1. / self.var_tau[c].astype(np.float64))
                for c, v in self.var_gamma.items()}

    def get_null_pi(self):                               # VIPRS.py:741-753
        return 1. - self.pi

    # VIPRS.py:497-581 (scalar pi / tau_beta branch)
    def elbo(self, sum_axis=None):
        res = np.finfo(np.float64).resolution
        var_gamma = np.clip(_dict_concat(self.var_gamma).astype(np.float64), a_min=res, a_max=1. - res)
        null_gamma = np.clip(1. - _dict_concat(self.compute_pip()).astype(np.float64), a_min=res, a_max=1. - res)
        log_var_tau = _dict_concat(self._log_var_tau)
        pi, null_pi, tau_beta = self.pi, self.get_null_pi(), self.tau_beta
        zeta = _dict_concat(self.zeta).astype(np.float64)
        elbo = 0.
        elbo -= np.log(2 * np.pi * self.sigma_epsilon)
        if "sigma_epsilon" not in self.fix_params:
            elbo -= 1.
        else:
            eta = _dict_concat(self.eta).astype(np.float64)
            std_beta = _dict_concat(self.std_beta).astype(np.float64)
            elbo -= (1. / self.sigma_epsilon) * (1. - 2. * std_beta.dot(eta) + self._sigma_g)
        elbo *= 0.5 * self.n
        elbo -= np.multiply(var_gamma, np.log(var_gamma) - np.log(pi)).sum(axis=sum_axis)
        elbo -= np.multiply(null_gamma, np.log(null_gamma) - np.log(null_pi)).sum(axis=sum_axis)
        elbo += .5 * np.multiply(var_gamma, 1. - log_var_tau + np.log(tau_beta)).sum(axis=sum_axis)
        elbo -= .5 * (tau_beta * zeta).sum(axis=sum_axis)
        return elbo

    def mse(self):                                       # VIPRS.py:689-704
        eta = _dict_concat(self.eta)
        std_beta = _dict_concat(self.std_beta)
        zeta = _dict_concat(self.zeta)
        return 1. - 2. * std_beta.dot(eta) + (self._sigma_g - zeta.sum(axis=None) + (eta ** 2).sum(axis=None))

    def get_heritability(self):                          # VIPRS.py:780-785
        return self._sigma_g / (self._sigma_g + self.sigma_epsilon)

    def run(self, n_iter, theta_0=None, param_0=None):
        """The body of VIPRS.fit's main loop (VIPRS.py:979-1000) for a fixed iteration count."""
        self.initialize(theta_0, param_0)
        for _ in range(n_iter):
            self.e_step()
            self.m_step()
            self.history["ELBO"].append(float(self.elbo()))
            self.history["sigma_epsilon"].append(float(self.sigma_epsilon))
            self.history["tau_beta"].append(np.array(self.tau_beta, dtype=np.float64).copy())
            self.history["pi"].append(np.array(self.pi, dtype=np.float64).copy())
            self.history["sigma_g"].append(float(self._sigma_g))
            self.history["max_eta_diff"].append(float(max(np.max(np.abs(d)) for d in self.eta_diff.values())))
            self.history["mse"].append(float(self.mse()))
        return self


class OracleVIPRSMix(OracleVIPRS):
    """Restatement of viprs/model/VIPRSMix.py (K slabs + null, (M,K) C-order arrays)."""

    def __init__(self, ld, std_beta, n_per_snp, K=1, prior_multipliers=None, **kw):
        super().__init__(ld, std_beta, n_per_snp, **kw)
        self.K = K
        if prior_multipliers is not None:
            self.d = np.array(prior_multipliers).astype(self.float_precision)
        else:
            self.d = 2 ** np.linspace(-min(K - 1, 7), 0, K).astype(self.float_precision)    # VIPRSMix.py:52
        self.n_per_snp2 = {c: n[:, None].astype(self.float_precision) for c, n in self.n_per_snp.items()}  # :56-59

    def _param_shapes(self):
        return {c: (m, self.K) for c, m in self.shapes.items()}

    def _n(self, c):
        return self.n_per_snp2[c]

    # VIPRSMix.py:61-167 (deterministic branches: "pis" or "pi" given together with sigma_epsilon)
    def initialize_theta(self, theta_0=None):
        theta_0 = dict(theta_0 or {})
        theta_0.update(self.fix_params)
        if "pis" in theta_0:
            self.pi = np.asarray(theta_0["pis"])
        else:
            self.pi = theta_0["pi"] * np.ones(self.K) / self.K
        self.sigma_epsilon = theta_0["sigma_epsilon"]
        if "tau_betas" in theta_0:
            self.tau_beta = np.asarray(theta_0["tau_betas"])
        elif "tau_beta" in theta_0:
            self.tau_beta = np.repeat(theta_0["tau_beta"], self.K)
        else:
            global_tau = self.n_snps * np.dot(1.0 / self.d, self.pi) / (1.0 - self.sigma_epsilon)   # :155-161
            self.tau_beta = self.d * global_tau
        ft = np.dtype(self.float_precision).type
        self.sigma_epsilon = ft(self.sigma_epsilon)
        self.pi = np.asarray(self.pi).astype(self.float_precision)
        self.lambda_min = ft(self.lambda_min)
        self._sigma_g = ft(0.0)

    # VIPRSMix.py:169-225
    def e_step(self):
        for c in self.shapes:
            tau_beta, pi = self.tau_beta, self.pi
            self.var_tau[c] = (self.n_per_snp2[c] * (1.0 + self.lambda_min) / self.sigma_epsilon) + tau_beta
            log_null_pi = np.ones_like(self.eta[c]) * np.log(1.0 - self.pi.sum())
            mu_mult = (self.n_per_snp2[c] / (self.var_tau[c] * self.sigma_epsilon)).astype(self.float_precision)
            u_logs = (np.log(pi) - np.log(1.0 - pi) + 0.5 * (np.log(tau_beta) - np.log(self.var_tau[c]))
                      ).astype(self.float_precision)
            e_step_mixture(self.ld_left_bound[c], self.ld_indptr[c], self.ld_data[c], self.std_beta[c],
                           self.var_gamma[c], self.var_mu[c], self.eta[c], self.q[c], self.eta_diff[c],
                           log_null_pi.astype(self.float_precision), np.ascontiguousarray(u_logs),
                           np.ascontiguousarray(np.sqrt(0.5 * self.var_tau[c]).astype(self.float_precision)),
                           np.ascontiguousarray(mu_mult), self.dequantize_scale, self.threads,
                           self.low_memory, kind=self.kind)
        self.zeta = self.compute_zeta()

    # VIPRSMix.py:227-260 + VIPRS.py:446-471
    def m_step(self):
        if "pis" not in self.fix_params:
            pi_estimate = _dict_sum(self.var_gamma, axis=0)
            if "pi" in self.fix_params:
                pi_estimate = self.fix_params["pi"] * pi_estimate / pi_estimate.sum()
            else:
                pi_estimate = pi_estimate / self.n_snps
            self.pi = pi_estimate
        if "tau_betas" not in self.fix_params:
            zetas = sum(self.compute_zeta(sum_axis=0).values())
            tau_beta_estimate = np.sum(self.pi) * self.n_snps / np.dot(self.d, zetas)
            tau_beta_estimate = self.d * tau_beta_estimate
            self.tau_beta = np.clip(tau_beta_estimate, a_min=1.0, a_max=None)
        self._sigma_g = np.sum([
            np.sum((1. + self.lambda_min) * self.zeta[c] + np.multiply(self.q[c], self.eta[c]), axis=0)
            for c in self.shapes.keys()], axis=0)
        if "sigma_epsilon" not in self.fix_params:
            sig_eps = 0.
            for c in self.shapes:
                sig_eps -= 2. * self.std_beta[c].dot(self.eta[c])
            self.sigma_epsilon = 1. + sig_eps + self._sigma_g

    def compute_pip(self):                               # VIPRSMix.py:297-301
        return {c: g.sum(axis=1) for c, g in self.var_gamma.items()}

    def compute_eta(self):                               # VIPRSMix.py:303-307
        return {c: (v * self.var_mu[c]).sum(axis=1) for c, v in self.var_gamma.items()}

    def compute_zeta(self, sum_axis=1):                  # VIPRSMix.py:309-316
        return {c: (v * (self.var_mu[c] ** 2 + (1.0 / self.var_tau[c]))).sum(axis=sum_axis)
                for c, v in self.var_gamma.items()}

    def get_null_pi(self):                               # VIPRSMix.py:262-274
        return 1.0 - np.sum(self.pi)

    # VIPRS.py:497-581 evaluated with (M,K) gammas: pi / tau_beta broadcast along K
    def elbo(self, sum_axis=None):
        res = np.finfo(np.float64).resolution
        var_gamma = np.clip(_dict_concat(self.var_gamma).astype(np.float64), a_min=res, a_max=1. - res)
        null_gamma = np.clip(1. - _dict_concat(self.compute_pip()).astype(np.float64), a_min=res, a_max=1. - res)
        # NOTE (reference quirk, kept on purpose): VIPRSMix.e_step never refreshes `_log_var_tau`
        # (VIPRSMix.py:187-204 takes np.log(var_tau) inline), so VIPRS.elbo (VIPRS.py:519) reads the
        # value cached at initialisation (VIPRS.py:359).
        log_var_tau = _dict_concat(self._log_var_tau)
        pi, null_pi, tau_beta = self.pi, self.get_null_pi(), self.tau_beta
        elbo = 0.
        elbo -= np.log(2 * np.pi * self.sigma_epsilon)
        if "sigma_epsilon" not in self.fix_params:
            elbo -= 1.
        else:
            eta = _dict_concat(self.eta).astype(np.float64)
            std_beta = _dict_concat(self.std_beta).astype(np.float64)
            elbo -= (1. / self.sigma_epsilon) * (1. - 2. * std_beta.dot(eta) + self._sigma_g)
        elbo *= 0.5 * self.n
        elbo -= np.multiply(var_gamma, np.log(var_gamma) - np.log(pi)).sum(axis=sum_axis)
        elbo -= np.multiply(null_gamma, np.log(null_gamma) - np.log(null_pi)).sum(axis=sum_axis)
        elbo += .5 * np.multiply(var_gamma, 1. - log_var_tau + np.log(tau_beta)).sum(axis=sum_axis)
        # zeta is (M,) and tau_beta is a (K,) array here => the `else` branch of VIPRS.py:568-573
        var_mu = _dict_concat(self.var_mu)
        var_tau = _dict_concat(self.var_tau)
        elbo -= .5 * (np.multiply(var_gamma, tau_beta) * (var_mu ** 2 + 1. / var_tau)).sum(axis=sum_axis)
        return elbo


# ---------------------------------------------------------------------------------------------
# numpy restatement of the reduced sums (include/viprs_b200.h VIPRS_B200_S_*): what m_step()/elbo()/mse() read
# from the per-SNP arrays (VIPRS.py:426-484, 497-581, 689-704; VIPRSMix.py:227-260), one row per model column /
# mixture component.  Checker for viprs_b200_sums_* and for the host M-step in viprs_b200/em_host.py.
# ---------------------------------------------------------------------------------------------
NSUMS = 16


def sums_numpy(var_gamma, var_mu, eta, q, eta_diff, std_beta, n_per_snp, theta, theta_logtau=None, q_scale=1.0,
               mixture=False):
    """var_gamma / var_mu: (M,), (M,G) [grid] or (M,K) [mixture=True]; theta: (ncol, 4) = sigma_eps, tau_beta, pi, lambda."""
    res = np.finfo(np.float64).resolution
    g = np.asarray(var_gamma, dtype=np.float64).reshape(len(std_beta), -1)
    mu = np.asarray(var_mu, dtype=np.float64).reshape(len(std_beta), -1)
    ncol = g.shape[1]
    theta = np.asarray(theta, dtype=np.float64).reshape(ncol, 4)
    tl = theta if theta_logtau is None else np.asarray(theta_logtau, dtype=np.float64).reshape(ncol, 4)
    n = np.asarray(n_per_snp, dtype=np.float64)[:, None]
    vt = n * (1. + theta[:, 3]) / theta[:, 0] + theta[:, 1]
    lvt = np.log(n * (1. + tl[:, 3]) / tl[:, 0] + tl[:, 1])
    gc = np.clip(g, res, 1. - res)
    out = np.zeros((ncol, NSUMS))
    out[:, 0] = g.sum(0)
    out[:, 1] = (g * mu * mu).sum(0)
    out[:, 8] = (g / vt).sum(0)
    out[:, 4] = (gc * np.log(gc)).sum(0)
    out[:, 10] = gc.sum(0)
    out[:, 9] = (gc * lvt).sum(0)
    out[:, 11] = (gc * (mu * mu + 1. / vt)).sum(0)
    beta = np.asarray(std_beta, dtype=np.float64)
    if mixture:
        pip = np.asarray(var_gamma).sum(axis=1).astype(np.float64)[:, None]
        et = np.asarray(eta, dtype=np.float64)[:, None]
        qq = np.asarray(q, dtype=np.float64)[:, None]
        dd = np.asarray(eta_diff, dtype=np.float64)[:, None]
        cols = slice(0, 1)
    else:
        pip, et = g, np.asarray(eta, dtype=np.float64).reshape(g.shape)
        qq, dd = np.asarray(q, dtype=np.float64).reshape(g.shape), np.asarray(eta_diff, dtype=np.float64).reshape(g.shape)
        cols = slice(0, ncol)
    ng = np.clip(1. - pip, res, 1. - res)
    out[cols, 2] = q_scale * (et * qq).sum(0)
    out[cols, 3] = (beta[:, None] * et).sum(0)
    out[cols, 5] = (ng * np.log(ng)).sum(0)
    out[cols, 12] = ng.sum(0)
    out[cols, 6] = (et * et).sum(0)
    out[cols, 7] = np.abs(dd).max(0) if len(beta) else 0.
    return out
