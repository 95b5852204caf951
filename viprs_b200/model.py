"""
Host-side mirror of the reference's model classes for the E-step path: ``VIPRS``, ``VIPRSMix`` and ``VIPRSGrid``
(/root/reference/viprs/model/VIPRS.py, VIPRSMix.py, gridsearch/VIPRSGrid.py) with the per-SNP state resident in HBM.

Same constructor keywords where they touch the path, same attribute names (``var_gamma, var_mu, var_tau, eta, zeta,
eta_diff, q, sigma_epsilon, tau_beta, pi, _sigma_g, history, optim_result``), same ``e_step() / m_step() / elbo() /
fit()`` semantics and stopping rules, same outputs (``pip, post_mean_beta, post_var_beta``).  What changes is where
the work happens: LD and all per-SNP arrays live on the GPU, one EM iteration is
``prepare -> sweep -> sums`` (three launches) plus a few hundred bytes over PCIe, and, when torch.distributed is
initialised and ``shard=True``, whole LD blocks are sharded across ranks with one small all-reduce per iteration.

magenpy is not a dependency: pass either a ``gdl`` exposing ``get_ld_matrices()`` / ``sumstats_table`` like
magenpy's GWADataLoader (VIPRS.py:153-172, BayesPRSModel.py:118-142), or ``data={chrom: dict(ld_data, ld_indptr,
ld_left_bound, std_beta, n_per_snp)}``.  There is no CPU fallback.
"""
import copy
import ctypes
import logging
import math

import numpy as np
import torch

from . import _lib, em_host
from .em_host import MixHyper, SlabHyper
from .e_step import e_step_device, e_step_grid_device, e_step_mixture_device, q_offset_device
from .ld import DeviceLD, _stream_ptr
from .optim import IterationConditionCounter, OptimizeResult
from .parallel import SumsExchange, shard_genome, slice_chromosome

logger = logging.getLogger(__name__)

_TORCH_FLOAT = {"float32": torch.float32, "float64": torch.float64}


def _chroms_from_gdl(gdl, low_memory, dequantize_on_the_fly, float_precision):
    """The arrays VIPRS.__init__ pulls out of a GWADataLoader (VIPRS.py:153-172; BayesPRSModel.py:118-142)."""
    out = {}
    for c, ld_mat in gdl.get_ld_matrices().items():
        stored = np.dtype(ld_mat.stored_dtype)
        dtype = stored if (dequantize_on_the_fly and np.issubdtype(stored, np.integer)) else float_precision
        lop = ld_mat.load(return_symmetric=not low_memory, dtype=dtype)
        ss = gdl.sumstats_table[c]
        out[c] = dict(ld_data=lop.ld_data, ld_indptr=lop.ld_indptr, ld_left_bound=lop.leftmost_idx,
                      std_beta=ss.get_snp_pseudo_corr(), n_per_snp=ss.n_per_snp)
    return out


class VIPRS:
    """Spike-and-slab VIPRS (one model).  See the module docstring for the mapping to the reference class."""

    _layout = 0          # (M, ncol) column-major
    _half_tau = 0        # sqrt(var_tau / 2) for cpp_e_step (e_step.hpp:404)

    def __init__(self, gdl=None, fix_params=None, tracked_params=None, lambda_min=None, float_precision="float32",
                 order="F", low_memory=True, dequantize_on_the_fly=False, threads=1, data=None, device=None,
                 shard=False, presharded=False, group=None):
        assert float_precision in _TORCH_FLOAT
        if not torch.cuda.is_available():
            raise _lib.ViprsB200Error(-5, "VIPRS (viprs_b200 has no CPU fallback)")
        self.float_precision = float_precision
        self._tdt = _TORCH_FLOAT[float_precision]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.gdl = gdl
        self.threads = threads            # accepted for API compatibility: the sweep is always the sequential order
        self.order = order
        self.low_memory = low_memory
        self.fix_params = dict(fix_params or {})
        self.tracked_params = list(tracked_params or [])
        if data is None:
            if gdl is None:
                raise ValueError("VIPRS needs a gdl or data={chrom: {...}}")
            data = _chroms_from_gdl(gdl, low_memory, dequantize_on_the_fly, float_precision)
        self.chromosomes = list(data.keys())
        first = data[self.chromosomes[0]]["ld_data"]
        first_dt = first.dtype
        is_int = first_dt in (torch.int8, torch.int16) if isinstance(first, torch.Tensor) else np.issubdtype(first_dt, np.integer)
        # VIPRS.py:203-207
        if is_int:
            imax = {1: 127, 2: 32767}[first.element_size() if isinstance(first, torch.Tensor) else first.dtype.itemsize]
            self.dequantize_on_the_fly = True
            self.dequantize_scale = 1.0 / imax
        else:
            self.dequantize_on_the_fly = False
            self.dequantize_scale = 1.0
        # VIPRS.py:177-196 (a vector-valued or 'infer' lambda_min is outside the path's scope)
        self.lambda_min = 0.0 if lambda_min is None else float(lambda_min)

        # ---- sharding: whole LD blocks per rank ----
        self.rank, self.world, self.group = 0, 1, group
        if (shard or presharded) and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank = torch.distributed.get_rank(group)
            self.world = torch.distributed.get_world_size(group)
        # global shapes (BayesPRSModel.py:60-75)
        self.shapes = {c: int(len(data[c]["std_beta"])) for c in self.chromosomes}
        self._n = float(max(float(torch.as_tensor(data[c]["n_per_snp"]).max()) for c in self.chromosomes))
        if self.world > 1 and presharded:
            # `data` is already this rank's shard (whole LD blocks): only the global sizes are exchanged
            import torch.distributed as dist
            sz = torch.tensor([float(self.shapes[c]) for c in self.chromosomes] + [0.0], dtype=torch.float64, device=self.device)
            dist.all_reduce(sz, op=dist.ReduceOp.SUM, group=group)
            mx = torch.tensor([self._n], dtype=torch.float64, device=self.device)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
            self.row_ranges = {c: (0, self.shapes[c]) for c in self.chromosomes}
            self.shapes = {c: int(v) for c, v in zip(self.chromosomes, sz.tolist())}
            self._n = float(mx.item())
        elif self.world > 1:
            plan = shard_genome({c: {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in data[c].items()
                                     if k in ("ld_left_bound", "ld_indptr")} for c in self.chromosomes}, self.world)
            self.row_ranges = plan[self.rank]
            local = {}
            for c in self.chromosomes:
                r0, r1 = self.row_ranges[c]
                ch = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in data[c].items()}
                local[c] = slice_chromosome(ch, r0, r1)
            data = local
        else:
            self.row_ranges = {c: (0, self.shapes[c]) for c in self.chromosomes}
        self._load(data)
        self.optim_result = OptimizeResult()
        self.history = {}
        self._sums = None
        self.pip = self.post_mean_beta = self.post_var_beta = None

    # ------------------------------------------------------------------------------------------
    # data
    # ------------------------------------------------------------------------------------------
    def _load(self, data):
        """Concatenate the (local) chromosomes into one device-resident genome: one sweep launch covers them all."""
        dev = self.device
        sizes = [int(len(data[c]["std_beta"])) for c in self.chromosomes]
        self.local_shapes = dict(zip(self.chromosomes, sizes))
        self._seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        self.M = int(self._seg[-1])
        to_t = lambda a, dt=None: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=dt)
        if self.M > 0:
            datas, ips, lbs = [], [], []
            off = 0
            for c, r0 in zip(self.chromosomes, self._seg[:-1]):
                d = data[c]
                ip = to_t(d["ld_indptr"], torch.int64)
                datas.append(to_t(d["ld_data"]))
                ips.append(ip[:-1] + off)
                off += int(ip[-1])
                lbs.append(to_t(d["ld_left_bound"], torch.int32) + int(r0))
            ips.append(torch.tensor([off], dtype=torch.int64, device=dev))
            ld_data, ld_indptr, ld_lb = torch.cat(datas), torch.cat(ips), torch.cat(lbs).to(torch.int32)
            with torch.cuda.device(dev):
                self.ld = DeviceLD(ld_data, ld_indptr, ld_lb)
            del ld_data, datas
            self.std_beta_dev = torch.cat([to_t(data[c]["std_beta"], self._tdt) for c in self.chromosomes]).contiguous()
            self.n_per_snp_dev = torch.cat([to_t(data[c]["n_per_snp"], torch.float64) for c in self.chromosomes]).contiguous()
        else:
            self.ld = None
            self.std_beta_dev = torch.zeros(0, dtype=self._tdt, device=dev)
            self.n_per_snp_dev = torch.zeros(0, dtype=torch.float64, device=dev)
        self._seg_dev = torch.from_numpy(self._seg).to(dev)

    @property
    def n_snps(self):
        return int(sum(self.shapes.values()))

    m = n_snps

    @property
    def n(self):
        return self._n

    def _views(self, t):
        """{chrom: view} of a concatenated per-SNP tensor ((M,), (M,K) row-major or (ncol, M) column-major storage)."""
        out = {}
        for c, r0, r1 in zip(self.chromosomes, self._seg[:-1], self._seg[1:]):
            out[c] = t[:, r0:r1].t() if (t.dim() == 2 and self._layout == 0) else t[r0:r1]
        return out

    # ------------------------------------------------------------------------------------------
    # initialisation (VIPRS.py:213-359)
    # ------------------------------------------------------------------------------------------
    @property
    def _ncol(self):
        return 1

    def initialize(self, theta_0=None, param_0=None):
        self.initialize_theta(theta_0)
        self.initialize_variational_parameters(param_0)
        self.init_optim_meta()
        self._upload_theta()          # elbo() / zeta / var_tau are usable straight after initialize(), as in the reference

    def init_optim_meta(self):
        self.history = {"ELBO": []}
        for tt in self.tracked_params:
            self.history[tt if isinstance(tt, str) else tt.__name__] = []
        self.optim_result.reset()

    def initialize_theta(self, theta_0=None):
        """VIPRS.py:245-316.  Random draws use numpy's global RNG like the reference."""
        theta_0 = dict(theta_0 or {})
        theta_0.update(self.fix_params)
        M = self.n_snps
        if "pi" not in theta_0:
            pi = np.random.uniform(low=max(10. / M, 1e-5), high=min(0.2, 1e4 / M))
        else:
            pi = theta_0["pi"]
        if "sigma_epsilon" not in theta_0:
            if "tau_beta" not in theta_0:
                # magenpy's simple_ldsc is not available here: the reference's own fallback (VIPRS.py:287-289)
                naive_h2g = np.random.uniform(low=.01, high=.1)
                se = 1. - naive_h2g
                tau = pi * M / max(naive_h2g, 0.01)
            else:
                tau = theta_0["tau_beta"]
                se = np.clip(1. - (pi * M / tau), a_min=1e-4, a_max=1. - 1e-4)
        else:
            se = theta_0["sigma_epsilon"]
            tau = theta_0["tau_beta"] if "tau_beta" in theta_0 else (pi * M) / np.maximum(0.01, 1. - se)
        if "lambda_min" in theta_0:                     # a grid over lambda_min reaches here through fix_params
            self.lambda_min = float(theta_0["lambda_min"])
        ft = np.dtype(self.float_precision).type        # the reference casts pi / sigma_epsilon to float_precision (:311-314)
        self._hyp = SlabHyper(float(ft(pi)), float(ft(se)), float(tau), float(ft(self.lambda_min)))
        self._sync_hyper()

    def _sync_hyper(self):
        """
        Sharded fit: the random draws above come from each rank's own numpy RNG, so rank 0's hyper-parameters are
        broadcast -- every shard must run the same model (the single-GPU model that rank 0 would have run).
        """
        if self.world <= 1:
            return
        import torch.distributed as dist
        h = self._hyp
        parts = [np.atleast_1d(np.asarray(getattr(h, k), dtype=np.float64)).ravel()
                 for k in ("pi", "sigma_epsilon", "tau_beta", "lambda_min")]
        buf = torch.from_numpy(np.concatenate(parts)).to(self.device)
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(buf, src=src, group=self.group)
        vals, o = buf.cpu().numpy(), 0
        for k, part in zip(("pi", "sigma_epsilon", "tau_beta", "lambda_min"), parts):
            v = vals[o:o + part.size]
            o += part.size
            cur = getattr(h, k)
            setattr(h, k, v.copy() if isinstance(cur, np.ndarray) else float(v[0]))

    def _alloc_state(self):
        dev, T, M, nc = self.device, self._tdt, self.M, self._ncol
        shape = (M,) if nc == 1 and self._layout == 0 else ((nc, M) if self._layout == 0 else (M, nc))
        z = lambda s=shape: torch.zeros(s, dtype=T, device=dev)
        self._g, self._mu = z(), z()
        vec = shape if self._layout == 0 else (M,)
        self._eta, self._q, self._diff = z(vec), z(vec), z(vec)
        self._ul, self._tt, self._mm = z(), z(), z()
        self._theta_dev = torch.zeros((nc, 4), dtype=torch.float64, device=dev)
        self._theta_host = torch.zeros((nc, 4), dtype=torch.float64).pin_memory()
        nseg = len(self.chromosomes)
        L = _lib.lib()
        wsb = int(L.viprs_b200_sums_workspace_bytes(max(M, 1), nc, nseg))
        self._ws = torch.zeros(max(wsb, 8), dtype=torch.uint8, device=dev)
        self._exchange = SumsExchange(nseg, nc, _lib.NSUMS, em_host.S_MAX_DIFF, self.rank, self.world, dev, self.group)
        # world > 1: the sums kernel writes straight into the collective's send buffer
        self._sums_dev = (self._exchange.table() if self.world > 1 else
                          torch.zeros((nseg, nc, _lib.NSUMS), dtype=torch.float64, device=dev))
        self._q_is_forward = False
        self._dev_em_ready = False          # device-resident EM buffers / the captured graph refer to the old arrays
        self._graph = None

    def initialize_variational_parameters(self, param_0=None):
        """VIPRS.py:318-359: mu = 0, gamma = pi, eta = gamma * mu, q = 0 (param_0 may override mu / gamma)."""
        param_0 = param_0 or {}
        self._alloc_state()
        self._set_gamma_to_pi()
        for name, t in (("mu", self._mu), ("gamma", self._g)):
            if name in param_0:
                for c, v in self._views(t).items():
                    r0, r1 = self.row_ranges[c]
                    v.copy_(torch.as_tensor(np.asarray(param_0[name][c])[r0:r1], dtype=self._tdt).to(self.device))
        self._eta.copy_(self._compute_eta())
        self._theta_logtau = self._hyp.theta().copy()
        self._theta_logtau[:, 3] = 0.0                   # VIPRS.py:329,359: n / sigma_epsilon + tau_beta, no lambda
        self._sums = None
        # The reference starts with q = 0 even when param_0 makes eta != 0 (VIPRS.py:355-357) and then maintains q
        # incrementally, so q carries the constant offset -dq (R - I) eta_0 for the whole fit.  The one-pass sweep
        # recomputes q from eta, so that offset is handed to it explicitly.
        self._qoff = None
        # (the batched grid sweep keeps q in/out incrementally, like the reference: nothing to do there)
        if "mu" in param_0 and self.M > 0 and (self._layout == 1 or self._ncol == 1):
            with torch.cuda.device(self.device):
                self._qoff = q_offset_device(self.ld, self._eta, self._q, self.dequantize_scale)

    def _set_gamma_to_pi(self):
        self._g.fill_(float(self._hyp.pi[0]))

    def _compute_eta(self):
        return self._g * self._mu

    # ------------------------------------------------------------------------------------------
    # the reference's attribute surface
    # ------------------------------------------------------------------------------------------
    var_gamma = property(lambda self: self._views(self._g))
    var_mu = property(lambda self: self._views(self._mu))
    eta = property(lambda self: self._views(self._eta))
    eta_diff = property(lambda self: self._views(self._diff))
    std_beta = property(lambda self: self._views(self.std_beta_dev))
    n_per_snp = property(lambda self: self._views(self.n_per_snp_dev))

    @property
    def q(self):
        """The reference's q (forward + backward part, e_step.hpp:435-440); materialised on demand."""
        self._materialize_q()
        return self._views(self._q)

    def _materialize_q(self):
        if self._q_is_forward and self.M > 0:
            self.ld.backward_dot(self._eta, self._q, self.dequantize_scale)
            self._q_is_forward = False

    @property
    def pi(self):
        return float(self._hyp.pi[0])

    @property
    def tau_beta(self):
        return float(self._hyp.tau_beta[0])

    @property
    def sigma_epsilon(self):
        return float(self._hyp.sigma_epsilon[0])

    @property
    def _sigma_g(self):
        return float(self._hyp.sigma_g[0])

    def _var_tau_full(self, theta):
        ov = getattr(self, "_var_tau_override", None)     # set by grid_post.bayesian_model_average (averaged var_tau)
        if ov is not None:
            return ov
        th = torch.as_tensor(theta, dtype=torch.float64, device=self.device)          # (ncol, 4)
        n = self.n_per_snp_dev
        vt = n[None, :] * ((1.0 + th[:, 3:4]) / th[:, 0:1]) + th[:, 1:2]              # (ncol, M)  VIPRS.py:400
        if self._layout == 1:
            return vt.t().contiguous()
        return vt[0] if self._ncol == 1 else vt

    @property
    def var_tau(self):
        return self._views(self._var_tau_full(self._theta_last))

    def compute_pip(self):                               # VIPRS.py:875-880
        return self.var_gamma

    def compute_eta(self):                               # VIPRS.py:882-886
        return self._views(self._compute_eta())

    def _zeta_full(self):
        g = self._g.to(torch.float64)
        return g * (self._mu.to(torch.float64) ** 2 + 1.0 / self._var_tau_full(self._theta_last))

    def compute_zeta(self):                              # VIPRS.py:888-897 (float64)
        return self._views(self._zeta_full())

    zeta = property(compute_zeta)

    def get_proportion_causal(self):
        return self.pi

    def get_heritability(self):                          # VIPRS.py:780-785
        return float(em_host.heritability(self._hyp.sigma_g[0], self._hyp.sigma_epsilon[0]))

    # ------------------------------------------------------------------------------------------
    # one EM iteration
    # ------------------------------------------------------------------------------------------
    def _upload_theta(self):
        th = self._hyp.theta()
        self._theta_last = th.copy()
        self._theta_host.copy_(torch.from_numpy(th))
        self._theta_dev.copy_(self._theta_host, non_blocking=True)

    def _prepare(self, upload=True):
        """VIPRS.py:400-406,418 on the device."""
        if upload:
            self._upload_theta()
        if self.M == 0:
            return
        L = _lib.lib()
        fn = L.viprs_b200_prepare_f32 if self._tdt == torch.float32 else L.viprs_b200_prepare_f64
        lnp = getattr(self, "_lnp", None)
        rc = fn(self.M, self._ncol, self._layout, self._half_tau, self.n_per_snp_dev.data_ptr(), self._theta_dev.data_ptr(),
                self._ul.data_ptr(), self._tt.data_ptr(), self._mm.data_ptr(), lnp.data_ptr() if lnp is not None else None,
                _stream_ptr())
        _lib.check(rc, "viprs_b200_prepare")

    def _sweep(self):
        self._sums_fused = False
        if (self._qoff is None and self._tdt == torch.float32 and self._ncol == 1 and self._layout == 0
                and getattr(self, "_fuse_ok", True)):
            # the M-step / ELBO reductions ride along in the sweep's output role (viprs_b200_e_step_fused_f32)
            L = _lib.lib()
            rc = L.viprs_b200_e_step_fused_f32(self.ld.handle, self.std_beta_dev.data_ptr(), self._g.data_ptr(),
                                               self._mu.data_ptr(), self._eta.data_ptr(), self._q.data_ptr(),
                                               self._diff.data_ptr(), self._ul.data_ptr(), self._tt.data_ptr(),
                                               self._mm.data_ptr(), float(self.dequantize_scale),
                                               self.n_per_snp_dev.data_ptr(), self._theta_dev.data_ptr(),
                                               len(self.chromosomes), self._seg_dev.data_ptr(), self._sums_dev.data_ptr(),
                                               _stream_ptr())
            if rc == 0:
                self._q_is_forward = True
                self._sums_fused = True
                return
            if rc != -6:                                   # -6: not covered by the fused kernel -> separate launches
                _lib.check(rc, "viprs_b200_e_step_fused")
            self._fuse_ok = False
        # with a q offset the sum eta'q cannot use the 2 x forward-part identity: materialise q every iteration
        e_step_device(self.ld, self.std_beta_dev, self._g, self._mu, self._eta, self._q, self._diff, self._ul, self._tt,
                      self._mm, self.dequantize_scale, self._qoff is not None, self._qoff)
        self._q_is_forward = self._qoff is None

    def e_step(self):
        """VIPRS.e_step (VIPRS.py:381-424): pre-compute + one Gauss-Seidel sweep over every LD block."""
        with torch.cuda.device(self.device):
            self._prepare()
            if self.M > 0:
                self._sweep()
        self._sums = None

    def _reduce(self):
        """The sums m_step() / elbo() / mse() need: one streaming kernel + (world > 1) one all-reduce."""
        if self._sums is not None:
            return self._sums
        with torch.cuda.device(self.device):
            if self.M > 0 and getattr(self, "_sums_fused", False):
                pass                                          # the sweep already left the table in _sums_dev
            elif self.M > 0:
                L = _lib.lib()
                fn = L.viprs_b200_sums_f32 if self._tdt == torch.float32 else L.viprs_b200_sums_f64
                tl = self._theta_logtau_dev()
                rc = fn(self.M, self._ncol, self._layout, len(self.chromosomes), self._seg_dev.data_ptr(),
                        self._g.data_ptr(), self._mu.data_ptr(), self._eta.data_ptr(), self._q.data_ptr(),
                        self._diff.data_ptr(), self.std_beta_dev.data_ptr(), self.n_per_snp_dev.data_ptr(),
                        self._theta_dev.data_ptr(), tl, 2.0 if self._q_is_forward else 1.0,
                        self._ws.data_ptr(), self._ws.numel(), self._sums_dev.data_ptr(), _stream_ptr())
                _lib.check(rc, "viprs_b200_sums")
            else:
                self._sums_dev.zero_()
            self._sums = self._exchange.all_reduce(self._sums_dev)
        return self._sums

    def _theta_logtau_dev(self):
        # VIPRS.e_step refreshes the log(var_tau) cache every iteration (VIPRS.py:401): same theta as the sweep (NULL)
        return None

    def _seg_sizes(self):
        return np.array([self.shapes[c] for c in self.chromosomes], dtype=np.float64)

    def m_step(self):
        """VIPRS.m_step (VIPRS.py:473-484)."""
        S = self._reduce()
        em_host.slab_m_step(S, self._seg_sizes(), self.n_snps, self._hyp, "pi" in self.fix_params,
                            "tau_beta" in self.fix_params, "sigma_epsilon" in self.fix_params)

    def elbo(self, sum_axis=None):
        """VIPRS.elbo (VIPRS.py:497-581)."""
        return float(em_host.slab_elbo(self._reduce(), self.n, self._hyp, "sigma_epsilon" in self.fix_params)[0])

    objective = elbo

    def mse(self):
        return float(em_host.slab_mse(self._reduce(), self._hyp)[0])

    def max_eta_diff(self):
        return float(em_host.max_eta_diff(self._reduce())[0])

    # ------------------------------------------------------------------------------------------
    # device-resident EM iterations: prepare -> sweep -> sums -> [all-reduce] -> scalar M-step / ELBO
    # with nothing crossing PCIe; the host looks at the per-iteration scalars only when it wants to
    # ------------------------------------------------------------------------------------------
    _HIST = 64           # device history ring (iterations)

    def _em_flags(self):
        f = np.zeros(self._ncol, dtype=np.int32)
        f += 1 * ("pi" in self.fix_params) + 2 * ("tau_beta" in self.fix_params) + 4 * ("sigma_epsilon" in self.fix_params)
        return f

    def _setup_device_em(self):
        dev, nc = self.device, self._ncol
        self._seg_sizes_dev = torch.from_numpy(self._seg_sizes()).to(dev)
        self._flags_dev = torch.from_numpy(self._em_flags()).to(dev)
        d = getattr(self, "d", None)
        self._mixd_dev = torch.from_numpy(np.asarray(d, dtype=np.float64)).to(dev) if (d is not None and self._layout == 1) else None
        self._theta_prev_dev = torch.zeros((nc, 4), dtype=torch.float64, device=dev)
        self._sigma_g_dev = torch.zeros(nc, dtype=torch.float64, device=dev)
        self._scal_dev = torch.zeros((self._HIST, nc, 8), dtype=torch.float64, device=dev)
        self._iter_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._graph = None
        self._dev_em_ready = True

    def _device_iteration(self):
        """One EM iteration, launches only (stream-ordered; capturable)."""
        L = _lib.lib()
        hook = getattr(self, "_iter_hook", None)          # (bench.py: L2 flush between iterations of small workloads)
        if hook is not None:
            hook()
        self._prepare(upload=False)
        if self.M > 0:
            self._sweep()
            if not getattr(self, "_sums_fused", False):
                fn = L.viprs_b200_sums_f32 if self._tdt == torch.float32 else L.viprs_b200_sums_f64
                rc = fn(self.M, self._ncol, self._layout, len(self.chromosomes), self._seg_dev.data_ptr(),
                        self._g.data_ptr(), self._mu.data_ptr(), self._eta.data_ptr(), self._q.data_ptr(),
                        self._diff.data_ptr(), self.std_beta_dev.data_ptr(), self.n_per_snp_dev.data_ptr(),
                        self._theta_dev.data_ptr(), self._theta_logtau_dev(), 2.0 if self._q_is_forward else 1.0,
                        self._ws.data_ptr(), self._ws.numel(), self._sums_dev.data_ptr(), _stream_ptr())
                _lib.check(rc, "viprs_b200_sums")
        else:
            self._sums_dev.zero_()
        table = self._exchange.reduce_on_device(self._sums_dev)
        onehot = self._exchange.onehot().data_ptr() if self.world > 1 else None
        fixpi = float(self.fix_params.get("pi", 0.0)) if self._layout == 1 else 0.0
        rc = L.viprs_b200_em_update(len(self.chromosomes), self._ncol, self._layout, self.world, table.data_ptr(), onehot,
                                    self._seg_sizes_dev.data_ptr(), self._flags_dev.data_ptr(),
                                    self._mixd_dev.data_ptr() if self._mixd_dev is not None else None,
                                    float(self.n_snps), float(self.n), fixpi, self._theta_dev.data_ptr(),
                                    self._theta_prev_dev.data_ptr(), self._sigma_g_dev.data_ptr(), self._scal_dev.data_ptr(),
                                    self._HIST, self._iter_dev.data_ptr(), _stream_ptr())
        _lib.check(rc, "viprs_b200_em_update")

    def em_iterations(self, k, graph=True):
        """
        Run ``k`` EM iterations entirely on the device and return their scalars as a (k, ncol, 8) float64 array
        (ELBO, mse, max |eta_diff|, h2, pi, tau_beta, sigma_epsilon, sigma_g per iteration and model column): ONE
        synchronisation and one small read-back for the k iterations.  ``graph=True`` replays the iteration as a CUDA
        graph (captured on first use; falls back to plain launches if capture is not possible).  The host-side
        hyper-parameters (``pi``, ``tau_beta``, ...) are refreshed from the device afterwards.
        """
        assert 0 < k <= self._HIST
        with torch.cuda.device(self.device):
            if not getattr(self, "_dev_em_ready", False):
                self._setup_device_em()
            self._flags_dev.copy_(torch.from_numpy(self._em_flags()))
            self._upload_theta()
            key = self._graph_key()
            if getattr(self, "_graph_key_captured", None) != key:
                self._graph, self._graph_key_captured = None, key      # the captured launches no longer describe this model
            it0 = int(self._iter_dev.item())
            remaining = k
            if graph and self._graph is None:
                self._device_iteration()                          # lazy initialisation happens outside the capture
                remaining -= 1
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._device_iteration()
                    self._graph = g
                except Exception as ex:                           # pragma: no cover
                    logger.warning(f"CUDA graph capture of the EM iteration failed ({ex!r}); using plain launches")
                    torch.cuda.synchronize()
                    self._graph = False
            for _ in range(remaining):
                if graph and self._graph:
                    self._graph.replay()
                else:
                    self._device_iteration()
            torch.cuda.synchronize()
            it1 = int(self._iter_dev.item())
            scal = self._scal_dev.cpu().numpy()
            hist = np.stack([scal[i % self._HIST] for i in range(it0, it1)]) if it1 > it0 else np.zeros((0, self._ncol, 8))
            self._pull_hyper()
        self._sums = None
        return hist

    def _graph_key(self):
        return (self._qoff is None,)

    def _pull_hyper(self):
        """Host copies of the hyper-parameters <- device theta (after device-resident iterations)."""
        th = self._theta_dev.cpu().numpy()
        self._theta_last = self._theta_prev_dev.cpu().numpy().copy()
        sg = self._sigma_g_dev.cpu().numpy()
        h = self._hyp
        if self._layout == 1:
            h.sigma_epsilon = float(th[0, 0]); h.tau_beta = th[:, 1].copy(); h.pi = th[:, 2].copy(); h.sigma_g = float(sg[0])
        else:
            h.sigma_epsilon[:] = th[:, 0]; h.tau_beta[:] = th[:, 1]; h.pi[:] = th[:, 2]; h.sigma_g[:] = sg

    def set_fixed_params(self, fix_params):
        """VIPRS.py:361-379."""
        self.fix_params.update(fix_params)
        ft = np.dtype(self.float_precision).type         # the reference casts every fixed value (:372-379)
        for key, val in fix_params.items():
            if key == "sigma_epsilon":
                self._hyp.sigma_epsilon[:] = float(ft(val))
            elif key == "tau_beta":
                self._hyp.tau_beta[:] = float(ft(val))
            elif key == "pi":
                self._hyp.pi[:] = float(ft(val))
            elif key == "lambda_min":
                self.lambda_min = float(ft(val))         # survives a re-initialisation (MSE restart) like the reference's
                self._hyp.lambda_min[:] = self.lambda_min

    def update_theta_history(self):
        """VIPRS.py:839-873 (the tracked quantities that exist on this path)."""
        self.history["ELBO"].append(self.elbo())
        for tt in self.tracked_params:
            if tt == "pi":
                self.history["pi"].append(self.get_proportion_causal())
            elif tt == "heritability":
                self.history["heritability"].append(self.get_heritability())
            elif tt == "sigma_epsilon":
                self.history["sigma_epsilon"].append(self.sigma_epsilon)
            elif tt == "tau_beta":
                self.history["tau_beta"].append(copy.copy(self.tau_beta))
            elif tt == "sigma_g":
                self.history["sigma_g"].append(self._sigma_g)
            elif tt == "mse":
                self.history["mse"].append(self.mse())
            elif tt == "max_eta_diff":
                self.history["max_eta_diff"].append(self.max_eta_diff())
            elif callable(tt):
                self.history[tt.__name__].append(tt(self))

    def update_posterior_moments(self):
        """VIPRS.py:899-907: numpy outputs on the host (this rank's rows)."""
        self._materialize_q()
        h = lambda d: {c: v.detach().cpu().numpy().copy() for c, v in d.items()}
        self.pip = h(self.compute_pip())
        self.post_mean_beta = h(self.eta)
        zeta, eta = self.compute_zeta(), self.eta
        self.post_var_beta = {c: (zeta[c] - eta[c].to(torch.float64) ** 2).cpu().numpy() for c in zeta}

    # ------------------------------------------------------------------------------------------
    # fit (VIPRS.py:909-1124)
    # ------------------------------------------------------------------------------------------
    def fit(self, max_iter=1000, theta_0=None, param_0=None, continued=False, disable_pbar=True, min_iter=3,
            f_abs_tol=1e-6, x_abs_tol=1e-6, patience=10, device_loop=False, check_every=8, **kwargs):
        """
        VIPRS.fit (VIPRS.py:909-1124): same initialisation, same per-iteration order (e_step, m_step, objective), same
        stopping rules and messages.

        ``device_loop=False`` (default): every iteration's sums come back to the host, which runs the scalar M-step and
        the checks -- the reference's control flow one to one.  ``device_loop=True``: iterations run on the device in
        chunks of ``check_every`` (CUDA graph of prepare -> sweep -> sums -> [all-reduce] -> scalar M-step / ELBO, see
        ``em_iterations``); the host applies the SAME rules to the recorded per-iteration scalars afterwards, so the
        history, the iteration at which the fit is declared finished and the message are the reference's, while the
        state may have advanced by up to ``check_every - 1`` further iterations (it only gets closer to the fixed point).
        """
        if not continued:
            self.initialize(theta_0, param_0)
            start_idx = 1
            self._upload_theta()
            self.update_theta_history()
            prev_elbo = -np.inf
        else:
            start_idx = len(self.history["ELBO"]) + 1
            self.optim_result.update(self.elbo(), increment=False)
            prev_elbo = self.elbo()
        prev_sigma_g = self._sigma_g
        sigma_g_icc, divergence_icc = IterationConditionCounter(), IterationConditionCounter()
        i = start_idx
        last = start_idx + max_iter
        pending = []                       # device loop: scalars of iterations already run but not yet examined
        while i < last:
            if self.optim_result.stop_iteration:
                break
            if device_loop:
                if not pending:
                    hist = self.em_iterations(min(max(int(check_every), 1), last - i))
                    pending = [dict(elbo=float(r[0, 0]) if self._layout == 0 else float(r[0, 0]), mse=float(r[0, 1]),
                                    max_eta_diff=float(r[0, 2]), h2=float(r[0, 3]), sigma_g=float(r[0, 7]),
                                    sigma_epsilon=float(r[0, 6]), pi=float(r[0, 4]), tau_beta=float(r[0, 5])) for r in hist]
                it = pending.pop(0)
                self.history["ELBO"].append(it["elbo"])
                for tt in self.tracked_params:
                    if isinstance(tt, str) and tt in it:
                        self.history[tt].append(it[tt])
                    elif tt == "heritability":
                        self.history["heritability"].append(it["h2"])
            else:
                self.e_step()
                self.m_step()
                self.update_theta_history()
                it = dict(elbo=self.history["ELBO"][-1], mse=self.mse(), max_eta_diff=self.max_eta_diff(),
                          h2=self.get_heritability(), sigma_g=self._sigma_g, sigma_epsilon=self.sigma_epsilon)
            max_eta_diff, curr_elbo, h2, mse, sg = it["max_eta_diff"], it["elbo"], it["h2"], it["mse"], it["sigma_g"]
            sigma_g_icc.update((i > min_iter) and np.isclose(sg, prev_sigma_g, atol=x_abs_tol, rtol=0.)
                               and max_eta_diff < x_abs_tol * 10, i)
            divergence_icc.update((curr_elbo < prev_elbo) and not np.isclose(curr_elbo, prev_elbo, atol=1e3 * f_abs_tol,
                                                                              rtol=1e-4), i)
            if mse < 0.:
                if "sigma_epsilon" not in self.fix_params:
                    logger.info(f"Iteration {i} | MSE is negative; restarting with sigma_epsilon fixed.")
                    self.initialize_theta(theta_0)
                    self.initialize_variational_parameters(param_0)
                    self.fix_params["sigma_epsilon"] = .95
                    if self._layout == 1:
                        self._hyp.sigma_epsilon = .95
                    else:
                        self._hyp.sigma_epsilon[:] = .95
                    pending = []
                    i += 1
                    continue
                self.optim_result.update(curr_elbo, stop_iteration=True, success=False,
                                         message=f"The MSE is negative ({mse:.6f}).")
            elif not np.isfinite(curr_elbo):
                self.optim_result.update(curr_elbo, stop_iteration=True, success=False, message="Objective (ELBO) is undefined.")
            elif it["sigma_epsilon"] < 0.:
                self.optim_result.update(curr_elbo, stop_iteration=True, success=False,
                                         message="Residual variance estimate is negative.")
            elif h2 > 1. or h2 < 0.:
                self.optim_result.update(curr_elbo, stop_iteration=True, success=False,
                                         message="Estimated heritability is out of bounds.")
            elif (i > min_iter) and np.isclose(prev_elbo, curr_elbo, atol=f_abs_tol, rtol=0.):
                self.optim_result.update(curr_elbo, stop_iteration=True, success=True,
                                         message="Objective (ELBO) converged successfully.")
            elif (i > min_iter) and max_eta_diff < x_abs_tol:
                self.optim_result.update(curr_elbo, stop_iteration=True, success=True,
                                         message="Variational parameters converged successfully.")
            elif sigma_g_icc.counter > patience:
                self.optim_result.update(curr_elbo, stop_iteration=True, success=True,
                                         message="LD-weighted variational parameters converged successfully.")
            elif divergence_icc.counter > patience:
                self.optim_result.update(curr_elbo, stop_iteration=True, success=False,
                                         message="The objective (ELBO) is decreasing.")
            else:
                self.optim_result.update(curr_elbo)
            prev_elbo = curr_elbo
            prev_sigma_g = sg
            i += 1
        self.update_posterior_moments()
        if not self.optim_result.stop_iteration:
            self.optim_result.update(self.history["ELBO"][-1] if device_loop else self.elbo(), stop_iteration=True, success=False,
                                     message="Maximum iterations reached without convergence.\n"
                                             "You may need to run the model for more iterations.", increment=False)
        if not self.optim_result.success:
            logger.warning("\t" + str(self.optim_result.message))
        return self


class VIPRSMix(VIPRS):
    """Sparse mixture prior, K slabs + null (VIPRSMix.py); (M,K) arrays are C-order."""

    _layout = 1

    def __init__(self, gdl=None, K=1, prior_multipliers=None, **kwargs):
        kwargs["order"] = "C"
        self.K = int(K)
        assert self.K > 0
        super().__init__(gdl, **kwargs)
        ft = np.dtype(self.float_precision)
        if prior_multipliers is not None:
            assert len(prior_multipliers) == self.K
            self.d = np.array(prior_multipliers).astype(ft)
        else:
            self.d = 2 ** np.linspace(-min(self.K - 1, 7), 0, self.K).astype(ft)        # VIPRSMix.py:52

    @property
    def _ncol(self):
        return self.K

    def initialize_theta(self, theta_0=None):
        """VIPRSMix.py:61-167."""
        theta_0 = dict(theta_0 or {})
        theta_0.update(self.fix_params)
        M, K, d = self.n_snps, self.K, self.d.astype(np.float64)
        if "pis" in theta_0:
            pis = np.asarray(theta_0["pis"], dtype=np.float64)
        else:
            overall = theta_0["pi"] if "pi" in theta_0 else np.random.uniform(low=max(0.005, 1.0 / M), high=0.1)
            pis = overall * np.random.dirichlet(np.ones(K))
        if "sigma_epsilon" not in theta_0:
            if "tau_betas" in theta_0:
                tau = np.asarray(theta_0["tau_betas"], dtype=np.float64)
                se = np.clip(1.0 - np.dot(1.0 / tau, pis), a_min=1e-4, a_max=1.0 - 1e-4)
            elif "tau_beta" in theta_0:
                tau = theta_0["tau_beta"] * d
                se = np.clip(1.0 - (M * pis / tau).sum(), a_min=1e-4, a_max=1.0 - 1e-4)
            else:
                naive_h2g = np.random.uniform(low=0.001, high=0.999)                     # ldsc unavailable: :137-138
                se = 1.0 - naive_h2g
                tau = d * (M * np.dot(1.0 / d, pis) / naive_h2g)
        else:
            se = theta_0["sigma_epsilon"]
            if "tau_betas" in theta_0:
                tau = np.asarray(theta_0["tau_betas"], dtype=np.float64)
            elif "tau_beta" in theta_0:
                tau = np.repeat(theta_0["tau_beta"], K).astype(np.float64)
            else:
                tau = d * (M * np.dot(1.0 / d, pis) / (1.0 - se))
        ft = np.dtype(self.float_precision).type
        self._hyp = MixHyper(np.asarray(pis).astype(ft).astype(np.float64), float(ft(se)), tau, d, float(ft(self.lambda_min)))
        self._sync_hyper()

    def _alloc_state(self):
        super()._alloc_state()
        self._lnp = torch.zeros(self.M, dtype=self._tdt, device=self.device)

    def _set_gamma_to_pi(self):
        self._g.copy_(torch.as_tensor(self._hyp.pi, dtype=self._tdt, device=self.device).expand(self.M, self.K))

    def _compute_eta(self):                              # VIPRSMix.py:303-307
        return (self._g * self._mu).sum(dim=1)

    def compute_pip(self):                               # VIPRSMix.py:297-301
        return self._views(self._g.sum(dim=1))

    def _zeta_full(self):                                # VIPRSMix.py:309-316 (sum over K)
        g = self._g.to(torch.float64)
        return (g * (self._mu.to(torch.float64) ** 2 + 1.0 / self._var_tau_full(self._theta_last))).sum(dim=1)

    pi = property(lambda self: self._hyp.pi.copy())
    tau_beta = property(lambda self: self._hyp.tau_beta.copy())
    sigma_epsilon = property(lambda self: float(self._hyp.sigma_epsilon))
    _sigma_g = property(lambda self: float(self._hyp.sigma_g))

    def get_proportion_causal(self):
        return float(np.sum(self._hyp.pi))

    def get_heritability(self):
        return float(em_host.heritability(self._hyp.sigma_g, self._hyp.sigma_epsilon))

    def _em_flags(self):
        f = np.zeros(self._ncol, dtype=np.int32)
        f[0] = (1 * ("pis" in self.fix_params) + 2 * ("tau_betas" in self.fix_params) + 4 * ("sigma_epsilon" in self.fix_params)
                + 8 * ("pi" in self.fix_params))
        return f

    def _sweep(self):
        self._sums_fused = False
        e_step_mixture_device(self.ld, self.std_beta_dev, self._g, self._mu, self._eta, self._q, self._diff, self._lnp,
                              self._ul, self._tt, self._mm, self.dequantize_scale, self._qoff is not None, self._qoff)
        self._q_is_forward = self._qoff is None

    def _theta_logtau_dev(self):
        # VIPRSMix.e_step never refreshes `_log_var_tau` (VIPRSMix.py:187-204 takes np.log(var_tau) inline), so the
        # reference's elbo() (VIPRS.py:519) keeps reading the value cached at initialisation (VIPRS.py:359).
        if getattr(self, "_tl_dev_src", None) is not self._theta_logtau:
            self._tl_dev = torch.from_numpy(self._theta_logtau).to(self.device)
            self._tl_dev_src = self._theta_logtau
        return self._tl_dev.data_ptr()

    def m_step(self):
        em_host.mix_m_step(self._reduce(), self.n_snps, self._hyp, self.fix_params)

    def elbo(self, sum_axis=None):
        return em_host.mix_elbo(self._reduce(), self.n, self._hyp, self.fix_params)

    objective = elbo

    def mse(self):
        return em_host.mix_mse(self._reduce(), self._hyp)

    def set_fixed_params(self, fix_params):
        self.fix_params.update(fix_params)
        for key, val in fix_params.items():
            if key == "sigma_epsilon":
                self._hyp.sigma_epsilon = float(val)
            elif key == "tau_betas":
                self._hyp.tau_beta = np.asarray(val, dtype=np.float64).copy()
            elif key == "pis":
                self._hyp.pi = np.asarray(val, dtype=np.float64).copy()
            elif key == "lambda_min":
                self.lambda_min = float(val)
                self._hyp.lambda_min = float(val)


class VIPRSGrid(VIPRS):
    """
    VIPRS over a grid of fixed hyper-parameters (gridsearch/VIPRSGrid.py).

    ``fit(pathwise=False)`` fits all grid points at once: the per-SNP state is an (M, n_models) column-major matrix,
    one ``e_step_grid`` sweep per EM iteration updates every column that is still iterating while the LD rows are read
    once (e_step.hpp:555-647), and every column runs the reference's own stopping rules; converged columns drop out
    of ``active_model_idx``.  ``fit(pathwise=True)`` (the reference's default: each model warm-started from the
    previous one, inherently sequential over grid points) loops the single-model path like the reference does.

    ``grid``: anything with ``to_table()`` returning a pandas DataFrame (HyperparameterGrid), a DataFrame, or a list
    of dicts with keys among ``pi``, ``sigma_epsilon``, ``tau_beta``, ``lambda_min``.
    """

    @property
    def _half_tau(self):
        # cpp_e_step_grid takes var_tau / 2 (e_step.hpp:616); the single-model sweep of the pathwise loop sqrt(var_tau / 2)
        return 1 if self._batched else 0

    def __init__(self, gdl=None, grid=None, **kwargs):
        if hasattr(grid, "to_table"):
            grid = grid.to_table()
        if hasattr(grid, "to_dict"):
            self.grid_table = grid
            records = grid.to_dict(orient="records")
        else:
            records = [dict(r) for r in grid]
            self.grid_table = records
        self._grid_records = records
        self.n_models = len(records)
        assert self.n_models > 1, "Grid search requires at least 2 models."
        self.validation_result = None
        self.optim_results = []
        self._batched = False
        super().__init__(gdl, **kwargs)

    @property
    def _ncol(self):
        return self.n_models if self._batched else 1

    models_to_keep = property(lambda self: np.logical_or(~self.terminated_models, self.converged_models))
    converged_models = property(lambda self: np.array([o.success for o in self.optim_results]))
    terminated_models = property(lambda self: np.array([o.stop_iteration for o in self.optim_results]))
    valid_terminated_models = property(lambda self: np.array([o.valid_optim_result for o in self.optim_results]))

    # ---- batched (independent columns) -------------------------------------------------------
    def _init_grid_hyper(self, theta_0):
        G, M = self.n_models, self.n_snps
        pi, se, tau, lam = np.empty(G), np.empty(G), np.empty(G), np.empty(G)
        self._fix = {k: np.zeros(G, dtype=bool) for k in ("pi", "sigma_epsilon", "tau_beta")}
        ft = np.dtype(self.float_precision).type
        for g, rec in enumerate(self._grid_records):
            saved_fix = self.fix_params
            self.fix_params = dict(saved_fix, **rec)
            VIPRS.initialize_theta(self, dict(theta_0 or {}))      # the single-model rules, per grid point
            for k in self._fix:
                self._fix[k][g] = k in self.fix_params
            if "lambda_min" in rec:
                self._hyp.lambda_min[:] = float(ft(rec["lambda_min"]))
            pi[g], se[g], tau[g], lam[g] = self._hyp.pi[0], self._hyp.sigma_epsilon[0], self._hyp.tau_beta[0], self._hyp.lambda_min[0]
            self.fix_params = saved_fix
        self._hyp = SlabHyper(pi, se, tau, lam)

    def _set_gamma_to_pi(self):
        if self._batched:
            self._g.copy_(torch.as_tensor(self._hyp.pi, dtype=self._tdt, device=self.device)[:, None].expand(self.n_models, self.M))
        else:
            super()._set_gamma_to_pi()

    def _sweep(self):
        if not self._batched:
            return super()._sweep()
        self._sums_fused = False
        key = tuple(self._active)
        if getattr(self, "_act_key", None) != key:                # (built outside any CUDA-graph capture)
            self._act_dev = torch.as_tensor(self._active, dtype=torch.int32, device=self.device)
            self._act_key = key
        act = self._act_dev
        if act.numel() == 0:
            return
        cm = lambda t: t.t()                              # (G, M) storage -> column-major (M, G) view
        e_step_grid_device(self.ld, self.std_beta_dev, cm(self._g), cm(self._mu), cm(self._eta), cm(self._q), cm(self._diff),
                           cm(self._ul), cm(self._tt), cm(self._mm), self.dequantize_scale, act)
        self._q_is_forward = False                        # the grid sweep keeps the reference's full q in place

    def _graph_key(self):
        return (self._batched, tuple(self._active) if self._batched else (), self._qoff is None)

    def _em_flags(self):
        if not self._batched:
            return super()._em_flags()
        return (1 * self._fix["pi"] + 2 * self._fix["tau_beta"] + 4 * self._fix["sigma_epsilon"]).astype(np.int32)

    pi = property(lambda self: self._hyp.pi.copy() if self._batched or self._hyp.ncol > 1 else float(self._hyp.pi[0]))
    tau_beta = property(lambda self: self._hyp.tau_beta.copy() if self._batched or self._hyp.ncol > 1 else float(self._hyp.tau_beta[0]))
    sigma_epsilon = property(lambda self: self._hyp.sigma_epsilon.copy() if self._batched or self._hyp.ncol > 1 else float(self._hyp.sigma_epsilon[0]))
    _sigma_g = property(lambda self: self._hyp.sigma_g.copy() if self._batched or self._hyp.ncol > 1 else float(self._hyp.sigma_g[0]))

    def get_heritability(self):
        h = em_host.heritability(self._hyp.sigma_g, self._hyp.sigma_epsilon)
        return h if (self._batched or self._hyp.ncol > 1) else float(h[0])

    def m_step(self):
        if not self._batched:
            return super().m_step()
        em_host.slab_m_step(self._reduce(), self._seg_sizes(), self.n_snps, self._hyp, self._fix["pi"], self._fix["tau_beta"],
                            self._fix["sigma_epsilon"], cols=self._active)

    def elbo(self, sum_axis=None):
        if not self._batched:
            return super().elbo()
        return em_host.slab_elbo(self._reduce(), self.n, self._hyp, self._fix["sigma_epsilon"])

    objective = elbo

    def mse(self):
        if not self._batched:
            return super().mse()
        return em_host.slab_mse(self._reduce(), self._hyp)

    def max_eta_diff(self):
        if not self._batched:
            return super().max_eta_diff()
        return em_host.max_eta_diff(self._reduce())

    def _fit_batched(self, max_iter=1000, theta_0=None, param_0=None, min_iter=3, f_abs_tol=1e-6, x_abs_tol=1e-6,
                     patience=10, **kwargs):
        G = self.n_models
        self._batched = True
        self._init_grid_hyper(theta_0)
        self.initialize_variational_parameters(param_0)
        self.history = {"ELBO": []}
        self.optim_results = [OptimizeResult() for _ in range(G)]
        for o in self.optim_results:
            o.reset()
        self._active = list(range(G))
        self._upload_theta()
        elbo = self.elbo()
        self.history["ELBO"].append(elbo.copy())
        prev_elbo = np.full(G, -np.inf)
        prev_sigma_g = self._hyp.sigma_g.copy()
        sg_icc = [IterationConditionCounter() for _ in range(G)]
        dv_icc = [IterationConditionCounter() for _ in range(G)]
        last_elbo = elbo.copy()
        for i in range(1, max_iter + 1):
            self._active = [g for g in range(G) if not self.optim_results[g].stop_iteration]
            if not self._active:
                break
            self.e_step()
            self.m_step()
            elbo_all, mse_all, med_all = self.elbo(), self.mse(), self.max_eta_diff()
            h2_all = em_host.heritability(self._hyp.sigma_g, self._hyp.sigma_epsilon)
            for g in self._active:
                o, curr, med = self.optim_results[g], float(elbo_all[g]), float(med_all[g])
                last_elbo[g] = curr
                sg_icc[g].update((i > min_iter) and np.isclose(self._hyp.sigma_g[g], prev_sigma_g[g], atol=x_abs_tol, rtol=0.)
                                 and med < x_abs_tol * 10, i)
                dv_icc[g].update((curr < prev_elbo[g]) and not np.isclose(curr, prev_elbo[g], atol=1e3 * f_abs_tol, rtol=1e-4), i)
                if mse_all[g] < 0. and not self._fix["sigma_epsilon"][g]:
                    # VIPRS.py:1025-1038: re-initialise this model and refit it with sigma_epsilon fixed at 0.95
                    logger.info(f"Iteration {i} | model {g}: MSE is negative; restarting with sigma_epsilon fixed.")
                    self._restart_column(g, theta_0, param_0)
                    continue
                if mse_all[g] < 0.:
                    o.update(curr, stop_iteration=True, success=False, message=f"The MSE is negative ({mse_all[g]:.6f}).")
                elif not np.isfinite(curr):
                    o.update(curr, stop_iteration=True, success=False, message="Objective (ELBO) is undefined.")
                elif self._hyp.sigma_epsilon[g] < 0.:
                    o.update(curr, stop_iteration=True, success=False, message="Residual variance estimate is negative.")
                elif h2_all[g] > 1. or h2_all[g] < 0.:
                    o.update(curr, stop_iteration=True, success=False, message="Estimated heritability is out of bounds.")
                elif (i > min_iter) and np.isclose(prev_elbo[g], curr, atol=f_abs_tol, rtol=0.):
                    o.update(curr, stop_iteration=True, success=True, message="Objective (ELBO) converged successfully.")
                elif (i > min_iter) and med < x_abs_tol:
                    o.update(curr, stop_iteration=True, success=True, message="Variational parameters converged successfully.")
                elif sg_icc[g].counter > patience:
                    o.update(curr, stop_iteration=True, success=True,
                             message="LD-weighted variational parameters converged successfully.")
                elif dv_icc[g].counter > patience:
                    o.update(curr, stop_iteration=True, success=False, message="The objective (ELBO) is decreasing.")
                else:
                    o.update(curr)
                prev_elbo[g] = curr
                prev_sigma_g[g] = self._hyp.sigma_g[g]
            self.history["ELBO"].append(last_elbo.copy())
        for g, o in enumerate(self.optim_results):
            if not o.stop_iteration:
                o.update(float(last_elbo[g]), stop_iteration=True, success=False,
                         message="Maximum iterations reached without convergence.\n"
                                 "You may need to run the model for more iterations.", increment=False)
        self.optim_result.nit = int(np.sum([o.nit for o in self.optim_results]))
        self._active = list(range(G))
        self.update_posterior_moments()
        self._finish_validation(last_elbo)
        return self

    def _restart_column(self, g, theta_0, param_0):
        """The reference's MSE-negative restart (VIPRS.py:1025-1038) for one column of the batched grid."""
        rec = dict(self._grid_records[g])
        saved_fix, saved_hyp = self.fix_params, self._hyp
        self.fix_params = dict(saved_fix, **rec)
        VIPRS.initialize_theta(self, dict(theta_0 or {}))
        one = self._hyp
        self.fix_params, self._hyp = saved_fix, saved_hyp
        self._hyp.pi[g], self._hyp.tau_beta[g], self._hyp.lambda_min[g] = one.pi[0], one.tau_beta[0], one.lambda_min[0]
        self._hyp.sigma_epsilon[g] = .95
        self._hyp.sigma_g[g] = 0.
        self._fix["sigma_epsilon"][g] = True
        self._mu[g].zero_(); self._q[g].zero_(); self._diff[g].zero_()
        self._g[g].fill_(float(self._hyp.pi[g]))
        for name, t in (("mu", self._mu), ("gamma", self._g)):
            if param_0 and name in param_0:
                for c, r0, r1 in zip(self.chromosomes, self._seg[:-1], self._seg[1:]):
                    a, b = self.row_ranges[c]
                    t[g, r0:r1].copy_(torch.as_tensor(np.asarray(param_0[name][c])[a:b], dtype=self._tdt).to(self.device))
        self._eta[g].copy_(self._g[g] * self._mu[g])
        self._theta_logtau[g] = [.95, self._hyp.tau_beta[g], self._hyp.pi[g], 0.0]
        self._sums = None

    def _finish_validation(self, elbos):
        self._elbo_final = np.asarray(elbos, dtype=np.float64).copy()      # per-model final ELBO (model selection / BMA)
        msgs = [o.message for o in self.optim_results]
        try:
            vr = self.grid_table.copy()
            vr["ELBO"] = elbos
            vr["Converged"] = self.converged_models
            vr["Optimization_message"] = msgs
        except Exception:
            vr = [dict(r, ELBO=float(e), Converged=bool(c), Optimization_message=m)
                  for r, e, c, m in zip(self._grid_records, elbos, self.converged_models, msgs)]
        self.validation_result = vr

    # ---- pathwise (the reference's serial warm-started loop, VIPRSGrid.py:176-248) -------------
    def _fit_pathwise(self, restart=False, **fit_kwargs):
        G = self.n_models
        self._batched = False
        cols = {k: [] for k in ("g", "mu", "q", "theta_last")}
        hyp = {k: np.empty(G) for k in ("sigma_epsilon", "pi", "sigma_g", "tau_beta")}
        elbos = np.empty(G)
        optim_results = []
        base_fix = dict(self.fix_params)
        for i, rec in enumerate(self._grid_records):
            # VIPRSGrid.py:197: set_fixed_params(params[i]) for EVERY i (lambda_min included); before the first fit
            # there are no hyper-parameters to overwrite yet, initialize_theta picks the values up from fix_params
            if hasattr(self, "_hyp") and self._hyp.ncol == 1:
                self.set_fixed_params(rec)
            else:
                self.fix_params.update(rec)
                if "lambda_min" in rec:
                    self.lambda_min = float(np.dtype(self.float_precision).type(rec["lambda_min"]))
            VIPRS.fit(self, continued=(i > 0 and not restart), **fit_kwargs)
            optim_results.append(copy.deepcopy(self.optim_result))
            self.optim_result.reset()
            elbos[i] = self.history["ELBO"][-1]
            self._materialize_q()
            cols["g"].append(self._g.clone()); cols["mu"].append(self._mu.clone()); cols["q"].append(self._q.clone())
            cols["theta_last"].append(self._theta_last[0].copy())      # var_tau of this model's last E-step (VIPRSGrid.py:219)
            hyp["sigma_epsilon"][i], hyp["pi"][i] = self._hyp.sigma_epsilon[0], self._hyp.pi[0]
            hyp["sigma_g"][i], hyp["tau_beta"][i] = self._hyp.sigma_g[0], self._hyp.tau_beta[0]
        self.fix_params = base_fix
        self.optim_results = optim_results
        self.optim_result.nit = int(np.sum([o.nit for o in optim_results]))
        # the model's attributes become (M, n_models) matrices (VIPRSGrid.py:233-248)
        lam = float(self._hyp.lambda_min[0])
        self._batched = True
        self._hyp = SlabHyper(hyp["pi"], hyp["sigma_epsilon"], hyp["tau_beta"], lam)
        self._hyp.sigma_g = hyp["sigma_g"]
        self._g, self._mu, self._q = torch.stack(cols["g"]), torch.stack(cols["mu"]), torch.stack(cols["q"])
        self._eta = self._g * self._mu
        self._q_is_forward = False
        self._theta_last = np.stack(cols["theta_last"])
        self._sums = None
        self._active = list(range(G))
        self.pip = {c: v.cpu().numpy().copy() for c, v in self.compute_pip().items()}
        self.post_mean_beta = {c: v.cpu().numpy().copy() for c, v in self.eta.items()}
        zeta, eta = self.compute_zeta(), self.eta
        self.post_var_beta = {c: (zeta[c] - eta[c].to(torch.float64) ** 2).cpu().numpy() for c in zeta}
        self._finish_validation(elbos)
        return self

    def select_best_model(self, criterion="ELBO", validation_std_beta=None):
        """grid_utils.select_best_model (grid_utils.py:8-123) on the device-resident matrices; see grid_post.py."""
        from . import grid_post
        return grid_post.select_best_model(self, criterion, validation_std_beta)

    def bayesian_model_average(self, normalization="softmax"):
        """grid_utils.bayesian_model_average (grid_utils.py:126-193) on the device-resident matrices; see grid_post.py."""
        from . import grid_post
        return grid_post.bayesian_model_average(self, normalization)

    def pseudo_validate(self, validation_std_beta):
        from . import grid_post
        return grid_post.pseudo_validate(self, validation_std_beta)

    def fit(self, pathwise=True, **fit_kwargs):
        self._var_tau_override = None
        fit_kwargs.pop("disable_pbar", None)
        if pathwise:
            return self._fit_pathwise(**fit_kwargs)
        return self._fit_batched(**fit_kwargs)
