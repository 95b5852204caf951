"""
The scalar side of one EM iteration: M-step, ELBO, MSE and heritability from the reduced per-chromosome sums the
device produces (``viprs_b200_sums_*``, slots ``VIPRS_B200_S_*`` of include/viprs_b200.h).  Pure numpy/float64, no
CUDA: the same code runs on every rank after the all-reduce, and in the CPU (gloo) tests.

Restates /root/reference/viprs/model/VIPRS.py:426-484 (m_step), 497-581 (elbo), 689-704 (mse), 780-785
(heritability) and VIPRSMix.py:227-260 on sums instead of length-M arrays.
"""
import numpy as np

NSUMS = 16
(S_GAMMA, S_GAMMA_MU2, S_ETA_Q, S_BETA_ETA, S_G_LOGG, S_NG_LOGNG, S_ETA2, S_MAX_DIFF, S_G_INV_TAU, S_G_LOG_TAU,
 S_GCLIP, S_GC_ZETA, S_NGCLIP) = range(13)


class SlabHyper:
    """Hyper-parameters of ncol independent spike-and-slab models (VIPRS: ncol = 1; VIPRSGrid: ncol = G)."""

    def __init__(self, pi, sigma_epsilon, tau_beta, lambda_min=0.0):
        self.pi = np.atleast_1d(np.asarray(pi, dtype=np.float64)).copy()
        self.sigma_epsilon = np.atleast_1d(np.asarray(sigma_epsilon, dtype=np.float64)).copy()
        self.tau_beta = np.atleast_1d(np.asarray(tau_beta, dtype=np.float64)).copy()
        n = self.pi.shape[0]
        self.lambda_min = np.broadcast_to(np.asarray(lambda_min, dtype=np.float64), (n,)).copy()
        self.sigma_g = np.zeros(n)

    @property
    def ncol(self):
        return self.pi.shape[0]

    def theta(self):
        """(ncol, 4) float64: sigma_epsilon, tau_beta, pi, lambda_min -- the device `theta` layout."""
        return np.ascontiguousarray(np.stack([self.sigma_epsilon, self.tau_beta, self.pi, self.lambda_min], axis=1))


def slab_m_step(S, seg_sizes, n_snps, hyp, fix_pi, fix_tau_beta, fix_sigma_epsilon, cols=None):
    """
    VIPRS.m_step (VIPRS.py:473-484) for every column in `cols` (default: all).  S: (nseg, ncol, NSUMS) global sums,
    seg_sizes: (nseg,) SNPs per chromosome.  fix_*: bool or (ncol,) bool arrays.  Updates `hyp` in place.
    """
    ncol = hyp.ncol
    cols = np.arange(ncol) if cols is None else np.asarray(cols)
    fp = np.broadcast_to(fix_pi, (ncol,))
    ft = np.broadcast_to(fix_tau_beta, (ncol,))
    fs = np.broadcast_to(fix_sigma_epsilon, (ncol,))
    seg_sizes = np.asarray(seg_sizes, dtype=np.float64)
    zeta = S[:, :, S_GAMMA_MU2] + S[:, :, S_G_INV_TAU]                   # (nseg, ncol) sum of zeta per chromosome
    # update_pi: dict_mean = unweighted mean over chromosomes of the per-chromosome mean (compute_utils.py:43-49)
    pi_new = np.mean(S[:, :, S_GAMMA] / seg_sizes[:, None], axis=0)
    for c in cols:
        if not fp[c]:
            hyp.pi[c] = pi_new[c]                                        # VIPRS.py:434
        if not ft[c]:
            hyp.tau_beta[c] = hyp.pi[c] * n_snps / zeta[:, c].sum()      # VIPRS.py:444
        hyp.sigma_g[c] = ((1.0 + hyp.lambda_min[c]) * zeta[:, c] + S[:, c, S_ETA_Q]).sum()      # VIPRS.py:454-457
        if not fs[c]:
            hyp.sigma_epsilon[c] = 1.0 - 2.0 * S[:, c, S_BETA_ETA].sum() + hyp.sigma_g[c]       # VIPRS.py:466-471
    return hyp


def slab_elbo(S, n, hyp, fix_sigma_epsilon, cols=None):
    """VIPRS.elbo (VIPRS.py:497-581), one value per column in `cols`; n = max n_per_snp (BayesPRSModel.py:75)."""
    ncol = hyp.ncol
    cols = np.arange(ncol) if cols is None else np.asarray(cols)
    fs = np.broadcast_to(fix_sigma_epsilon, (ncol,))
    T = S.sum(axis=0)                                                    # (ncol, NSUMS)
    out = np.empty(len(cols))
    with np.errstate(all="ignore"):
        for i, c in enumerate(cols):
            se, pi, tau = hyp.sigma_epsilon[c], hyp.pi[c], hyp.tau_beta[c]
            e = -np.log(2.0 * np.pi * se)                                # :545
            if not fs[c]:
                e -= 1.0                                                 # :552
            else:
                e -= (1.0 / se) * (1.0 - 2.0 * T[c, S_BETA_ETA] + hyp.sigma_g[c])          # :558
            e *= 0.5 * n                                                 # :560
            e -= T[c, S_G_LOGG] - np.log(pi) * T[c, S_GCLIP]             # :562
            e -= T[c, S_NG_LOGNG] - np.log(1.0 - pi) * T[c, S_NGCLIP]    # :563
            e += 0.5 * (T[c, S_GCLIP] * (1.0 + np.log(tau)) - T[c, S_G_LOG_TAU])           # :565
            e -= 0.5 * tau * (T[c, S_GAMMA_MU2] + T[c, S_G_INV_TAU])     # :568
            out[i] = e
    return out


def slab_mse(S, hyp, cols=None):
    """VIPRS.mse (VIPRS.py:689-704) per column."""
    cols = np.arange(hyp.ncol) if cols is None else np.asarray(cols)
    T = S.sum(axis=0)
    zeta = T[:, S_GAMMA_MU2] + T[:, S_G_INV_TAU]
    return np.array([1.0 - 2.0 * T[c, S_BETA_ETA] + (hyp.sigma_g[c] - zeta[c] + T[c, S_ETA2]) for c in cols])


def max_eta_diff(S, cols=None):
    """max |eta_diff| over all chromosomes (VIPRS.py:997), per column."""
    m = S[:, :, S_MAX_DIFF].max(axis=0)
    return m if cols is None else m[np.asarray(cols)]


def heritability(sigma_g, sigma_epsilon):
    """VIPRS.get_heritability (VIPRS.py:780-785)."""
    with np.errstate(all="ignore"):
        return sigma_g / (sigma_g + sigma_epsilon)


class MixHyper:
    """Hyper-parameters of the sparse mixture (VIPRSMix): pi / tau_beta are (K,) arrays."""

    def __init__(self, pis, sigma_epsilon, tau_betas, d, lambda_min=0.0):
        self.pi = np.asarray(pis, dtype=np.float64).copy()
        self.tau_beta = np.asarray(tau_betas, dtype=np.float64).copy()
        self.sigma_epsilon = float(sigma_epsilon)
        self.lambda_min = float(lambda_min)
        self.d = np.asarray(d, dtype=np.float64).copy()
        self.sigma_g = 0.0

    @property
    def ncol(self):
        return self.pi.shape[0]

    def theta(self):
        K = self.ncol
        return np.ascontiguousarray(np.stack([np.full(K, self.sigma_epsilon), self.tau_beta, self.pi,
                                              np.full(K, self.lambda_min)], axis=1))


def mix_m_step(S, n_snps, hyp, fix_params):
    """VIPRSMix.update_pi / update_tau_beta (VIPRSMix.py:227-260) + VIPRS._update_sigma_g / update_sigma_epsilon."""
    T = S.sum(axis=0)                                                    # (K, NSUMS)
    if "pis" not in fix_params:
        est = T[:, S_GAMMA].copy()                                       # dict_sum(var_gamma, axis=0)
        if "pi" in fix_params:
            est = fix_params["pi"] * est / est.sum()                     # :237
        else:
            est = est / n_snps                                           # :239
        hyp.pi = est
    zetas = T[:, S_GAMMA_MU2] + T[:, S_G_INV_TAU]                        # compute_zeta(sum_axis=0), summed over chromosomes
    if "tau_betas" not in fix_params:
        t = np.sum(hyp.pi) * n_snps / np.dot(hyp.d, zetas)               # :257
        hyp.tau_beta = np.clip(hyp.d * t, a_min=1.0, a_max=None)         # :258-260
    hyp.sigma_g = float((1.0 + hyp.lambda_min) * zetas.sum() + T[0, S_ETA_Q])                   # VIPRS.py:454-457
    if "sigma_epsilon" not in fix_params:
        hyp.sigma_epsilon = float(1.0 - 2.0 * T[0, S_BETA_ETA] + hyp.sigma_g)                   # VIPRS.py:466-471
    return hyp


def mix_elbo(S, n, hyp, fix_params):
    """VIPRS.elbo evaluated with (M,K) gammas (VIPRS.py:497-581; the `else` branch of :568-573)."""
    T = S.sum(axis=0)
    se = hyp.sigma_epsilon
    with np.errstate(all="ignore"):
        e = -np.log(2.0 * np.pi * se)
        if "sigma_epsilon" not in fix_params:
            e -= 1.0
        else:
            e -= (1.0 / se) * (1.0 - 2.0 * T[0, S_BETA_ETA] + hyp.sigma_g)
        e *= 0.5 * n
        e -= np.sum(T[:, S_G_LOGG] - np.log(hyp.pi) * T[:, S_GCLIP])
        e -= T[0, S_NG_LOGNG] - np.log(1.0 - np.sum(hyp.pi)) * T[0, S_NGCLIP]
        e += 0.5 * np.sum(T[:, S_GCLIP] * (1.0 + np.log(hyp.tau_beta)) - T[:, S_G_LOG_TAU])
        e -= 0.5 * np.sum(hyp.tau_beta * T[:, S_GC_ZETA])
    return float(e)


def mix_mse(S, hyp):
    T = S.sum(axis=0)
    zeta = (T[:, S_GAMMA_MU2] + T[:, S_G_INV_TAU]).sum()
    return float(1.0 - 2.0 * T[0, S_BETA_ETA] + (hyp.sigma_g - zeta + T[0, S_ETA2]))
