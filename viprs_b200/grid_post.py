"""
What happens to a fitted grid right after the E-step path: model selection and Bayesian model averaging over the
``(M, n_models)`` matrices (/root/reference/viprs/model/gridsearch/grid_utils.py:8-193) and the pseudo-validation
metric they may use (BayesPRSModel.py:385-410, eval/pseudo_metrics.py:130-152).  The matrices stay on the device: each
function is a handful of column reductions over the resident state, and the model is collapsed to a single column in
place -- afterwards it behaves like a fitted ``VIPRS`` (``pip / post_mean_beta / post_var_beta`` refreshed).

Only the criteria that live on this path are here: ``ELBO`` and ``pseudo_validation`` with standardized betas of a
validation set matched to the model's SNPs.  ``validation`` (individual-level genotypes, ``model.predict``) belongs to
magenpy's scoring code and is out of scope.
"""
import logging

import numpy as np
import torch

from .em_host import SlabHyper

logger = logging.getLogger(__name__)


def _valid(model):
    v = np.asarray(model.valid_terminated_models, dtype=bool)
    if v.shape[0] != model.n_models:
        raise ValueError("the grid has not been fitted")
    return v


def _final_elbos(model):
    """Final ELBO of every grid column, as fit() recorded it (VIPRSGrid.py:208, the `ELBO` column of validation_result)."""
    e = getattr(model, "_elbo_final", None)
    if e is None:
        raise ValueError("the grid has not been fitted")
    return np.asarray(e, dtype=np.float64).reshape(-1)


def pseudo_validate(model, validation_std_beta):
    """
    Pseudo-R^2 of every grid column against standardized marginal betas of a validation set
    (eval/pseudo_metrics.py:149-152 with the LD-weighted effects ``q + eta``, BayesPRSModel.py:397-399):
    ``(sum_j beta_val_j eta_j)^2 / sum_j eta_j (q_j + eta_j)``.  ``validation_std_beta``: {chrom: array} matched to the
    model's SNPs (this rank's rows when sharded), or one concatenated array.
    """
    if isinstance(validation_std_beta, dict):
        parts = []
        for c in model.chromosomes:
            r0, r1 = model.row_ranges[c]
            parts.append(np.asarray(validation_std_beta[c], dtype=np.float64)[r0:r1])
        vb = np.concatenate(parts)
    else:
        vb = np.asarray(validation_std_beta, dtype=np.float64)
    vb = torch.from_numpy(vb).to(model.device)
    model._materialize_q()
    eta = model._eta.to(torch.float64).reshape(-1, model.M)            # (n_models, M) storage
    q = model._q.to(torch.float64).reshape(-1, model.M)
    stats = torch.stack([(eta * vb[None, :]).sum(dim=1), (eta * (q + eta)).sum(dim=1)])
    if model.world > 1:
        import torch.distributed as dist
        dist.all_reduce(stats, group=model.group)
    rb, bsb = stats.cpu().numpy()
    with np.errstate(all="ignore"):
        return rb ** 2 / bsb


def _collapse(model, g_col, mu_col, q_col, theta_last_row, hyp):
    """The grid model becomes a single fitted model (grid_utils.py:71-123 / 168-191)."""
    model._batched = False
    model._g, model._mu, model._q = g_col.contiguous(), mu_col.contiguous(), q_col.contiguous()
    model._eta = model._g * model._mu
    model._diff = torch.zeros_like(model._g)
    model._ul, model._tt, model._mm = (torch.zeros_like(model._g) for _ in range(3))
    model._hyp = hyp
    model._theta_last = np.asarray(theta_last_row, dtype=np.float64).reshape(1, 4)
    model._theta_logtau = model._theta_last.copy()
    model._q_is_forward = False
    model._qoff = None
    model._sums = None
    model._sums_fused = False
    model._dev_em_ready = False
    model._graph = None
    model.n_models = 1
    nseg = len(model.chromosomes)
    from . import _lib, em_host
    from .parallel import SumsExchange
    wsb = int(_lib.lib().viprs_b200_sums_workspace_bytes(max(model.M, 1), 1, nseg))
    model._ws = torch.zeros(max(wsb, 8), dtype=torch.uint8, device=model.device)
    model._theta_dev = torch.zeros((1, 4), dtype=torch.float64, device=model.device)
    model._theta_host = torch.zeros((1, 4), dtype=torch.float64).pin_memory()
    model._exchange = SumsExchange(nseg, 1, _lib.NSUMS, em_host.S_MAX_DIFF, model.rank, model.world, model.device, model.group)
    model._sums_dev = (model._exchange.table() if model.world > 1 else
                       torch.zeros((nseg, 1, _lib.NSUMS), dtype=torch.float64, device=model.device))
    th = model._theta_last
    model._theta_host.copy_(torch.from_numpy(th))
    model._theta_dev.copy_(model._theta_host)
    model.update_posterior_moments()
    return model


def select_best_model(model, criterion="ELBO", validation_std_beta=None):
    """
    grid_utils.select_best_model (grid_utils.py:8-123) for ``criterion`` in (``ELBO``, ``pseudo_validation``): the best
    column among the validly terminated ones survives; ties and non-finite scores follow the reference (argmax of the
    scores with invalid models at -inf; pseudo-validation scores pass through nan_to_num first).
    """
    assert criterion in ("ELBO", "pseudo_validation")
    ok = _valid(model)
    if ok.sum() < 2:
        raise ValueError("Less than two models converged successfully. Cannot perform model selection.")
    if criterion == "ELBO":
        score = _final_elbos(model).copy()
        score[~ok] = -np.inf
        best = int(np.argmax(score))
    else:
        if validation_std_beta is None:
            raise ValueError("pseudo_validation needs the standardized betas of a validation set")
        score = pseudo_validate(model, validation_std_beta)
        score[~ok] = -np.inf
        try:
            model.validation_result["Pseudo_Validation_R2"] = score
        except Exception:
            pass
        best = int(np.argmax(np.nan_to_num(score, nan=0., neginf=0., posinf=0.)))
    logger.info(f"> Based on the {criterion} criterion, selected model: {best}")
    model._materialize_q()
    h = model._hyp
    hyp = SlabHyper(h.pi[best], h.sigma_epsilon[best], h.tau_beta[best], h.lambda_min[best])
    hyp.sigma_g = np.array([h.sigma_g[best]])
    rec = dict(model._grid_records[best])
    _collapse(model, model._g[best], model._mu[best], model._q[best], model._theta_last[best], hyp)
    model.best_model_idx = best
    model.set_fixed_params(rec)
    return model


def bayesian_model_average(model, normalization="softmax"):
    """
    grid_utils.bayesian_model_average (grid_utils.py:126-193): var_gamma, var_mu, var_tau and q become ELBO-weighted
    averages over the validly terminated columns (weights: softmax of the final ELBOs, or the shifted ELBOs normalised
    to one), eta / zeta follow, and the hyper-parameters are re-estimated with one unconstrained M-step on the averaged
    state.  The weights are taken over the kept columns (the reference computes them over all columns, which only
    type-checks when every model is kept).
    """
    if model.n_models < 2:
        return model
    ok = _valid(model)
    if ok.sum() < 1:
        raise ValueError("No models converged successfully. Cannot average models.")
    keep = np.flatnonzero(ok)
    elbos = _final_elbos(model)[keep]
    if normalization == "softmax":
        w = np.exp(elbos - elbos.max())
        w /= w.sum()
    elif normalization == "sum":
        w = elbos - elbos.min() + 1.
        w /= w.sum()
    else:
        raise KeyError(f"Normalization scheme not recognized. Valid options are: `softmax`, `sum`. Got: {normalization}")
    model._materialize_q()
    dev = model.device
    wt = torch.from_numpy(w).to(dev)
    idx = torch.from_numpy(keep).to(dev)
    avg = lambda t: (t.index_select(0, idx).to(torch.float64) * wt[:, None]).sum(dim=0).to(t.dtype)
    g, mu, q = avg(model._g), avg(model._mu), avg(model._q)
    # var_tau is averaged like the other matrices (grid_utils.py:168-171); it enters zeta = gamma (mu^2 + 1 / var_tau)
    vt = (model._var_tau_full(model._theta_last).index_select(0, idx) * wt[:, None]).sum(dim=0)        # (M,) float64
    h = model._hyp
    hyp = SlabHyper(float(np.dot(w, h.pi[keep])), float(np.dot(w, h.sigma_epsilon[keep])), float(np.dot(w, h.tau_beta[keep])),
                    float(h.lambda_min[keep[0]]))
    _collapse(model, g, mu, q, [hyp.sigma_epsilon[0], hyp.tau_beta[0], hyp.pi[0], hyp.lambda_min[0]], hyp)
    # the averaged var_tau is not n / sigma_epsilon + tau_beta of any single theta: keep it as an explicit override
    model._var_tau_override = vt
    model.update_posterior_moments()
    # re-estimate the hyper-parameters from the averaged state, nothing fixed (grid_utils.py:181-186 -> VIPRS.m_step,
    # VIPRS.py:426-484, with zeta from the averaged var_tau): a handful of float64 reductions over the resident arrays
    g64, mu64, q64 = g.to(torch.float64), mu.to(torch.float64), q.to(torch.float64)
    eta64 = (g * mu).to(torch.float64)
    zeta = g64 * (mu64 ** 2 + 1.0 / vt)
    lam = float(hyp.lambda_min[0])
    seg = model._seg
    per = []
    for r0, r1 in zip(seg[:-1], seg[1:]):
        per.append(torch.stack([g64[r0:r1].sum(), zeta[r0:r1].sum(), (q64[r0:r1] * eta64[r0:r1]).sum(),
                                (model.std_beta_dev[r0:r1].to(torch.float64) * eta64[r0:r1]).sum()]))
    S = torch.stack(per)                                                  # (nseg, 4)
    if model.world > 1:
        import torch.distributed as dist
        dist.all_reduce(S, group=model.group)
    S = S.cpu().numpy()
    sizes = model._seg_sizes()
    hyp.pi[0] = float(np.mean(S[:, 0] / sizes))                           # :434, dict_mean of per-chromosome means
    hyp.tau_beta[0] = hyp.pi[0] * model.n_snps / S[:, 1].sum()            # :444
    hyp.sigma_g[0] = ((1.0 + lam) * S[:, 1] + S[:, 2]).sum()              # :454-457
    hyp.sigma_epsilon[0] = 1.0 - 2.0 * S[:, 3].sum() + hyp.sigma_g[0]     # :466-471
    model.bma_weights = dict(zip(keep.tolist(), w.tolist()))
    return model
