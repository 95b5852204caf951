#!/usr/bin/env python
"""
bench.py -- E-step SNP-updates/s of the B200 coordinate-ascent sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c1|c5|ln|small]
                  [--scaling strong|weak] [--no-extras] [--no-e2e] [--no-cpu-baseline]

One "step" = the E-step of one EM iteration over the genome, exactly as `fit()` runs it: the per-SNP pre-compute, the
Gauss-Seidel sweep over every LD block, the M-step / ELBO reductions, the read-back of those few hundred bytes and the
scalar M-step; with N > 1 also the one small NCCL all-reduce.  Workloads (synthetic block-diagonal
PD LD, per-block Philox streams keyed on (7209, block id) -- any subset of blocks reproduces the full genome's arrays):
  c2 (default, BASELINE.json configs[1]): VIPRS spike-and-slab, 1,101,824 SNPs = 269 LD blocks x 4096, int8 LD
      (2047.5 stored entries per row, upper-triangular), float32 state, G = 1.
  c3: VIPRSGrid, same LD, 256 (pi x sigma_epsilon) grid columns sharing every LD row (e_step_grid semantics).
  c4: VIPRSMix K = 4, int16 LD.       c1: chr22-shaped 15,935 SNPs, float32 LD.       small: 16 blocks of c2.
  ln: c2 with LDetect-like log-normal block sizes (median ~650 SNPs, ~1,370 blocks).
  c5: BASELINE configs[4]: float64 state + float64 LD, 10,240-SNP blocks (~5,120 stored entries per row, tiled sweep);
      586 blocks = 6.0 M SNPs over 8 GPUs -- with N < 8 ranks the first 73 N blocks (30 GB of LD per GPU).
Multi-GPU is STRONG scaling by default: the fixed genome's LD blocks are dealt to the ranks (contiguous runs balanced
by sweep cost); the hyper-parameters are global (one all-reduce per step).  `--scaling weak` gives every rank its own
genome-sized shard instead.

Keys of the JSON line (rank 0):
  value      SNP-updates/s (SNPs x grid columns x steps / device time, max over ranks), everything resident in HBM.
  roofline   the sweep kernel alone: algorithmic bytes per launch / mean launch duration (CUDA events
             around the launch, inside the timed region) / measured HBM copy peak (MEASURED_PEAKS.json).  For c3 the
             binding resource is the CUDA-core FP32 pipe, reported as fp32_pipe_frac next to the HBM fraction.
  e2e        ONE C-ABI call per step with HOST (pinned) state arrays on the resident LD matrix
             (viprs_b200_cpp_e_step_resident = the argument list of cpp_e_step, q included): host->device of what
             cpp_e_step reads, sweep, q materialised, device->host of what it writes, all inside the timed region.
  workloads  device-timed c3 / c4 / c1 next to the headline c2 (same sharding), so that one driver run records them.
  cpu_baseline  the reference's own C++ e_step (compiled unmodified, oracle/_ref) on the host cores over the first 16
             LD blocks of the same arrays: all threads (the racy "hogwild" variant) and threads = 1 (the parity order).
"""
import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(model="viprs", n_blocks=269, block=4096, ld_dtype="int8", G=1, K=1, default_steps=300, fp="float32",
               desc="VIPRS spike-and-slab, 1,101,824 SNPs = 269 LD blocks x 4096, int8 LD (2047.5 nnz/row), float32, G=1"),
    "c3": dict(model="grid", n_blocks=269, block=4096, ld_dtype="int8", G=256, K=1, default_steps=5, fp="float32",
               desc="VIPRSGrid 256 (16 pi x 16 sigma_epsilon) columns sharing LD rows, 1,101,824 SNPs = 269 LD blocks x 4096, "
                    "int8 LD, float32"),
    "c4": dict(model="mix", n_blocks=269, block=4096, ld_dtype="int16", G=1, K=4, default_steps=100, fp="float32",
               desc="VIPRSMix K=4, 1,101,824 SNPs = 269 LD blocks x 4096, int16 LD, float32"),
    "c1": dict(model="viprs", n_blocks=0, block=0, ld_dtype="float32", G=1, K=1, M=15935, default_steps=300, fp="float32",
               desc="VIPRS spike-and-slab, chr22-shaped 15,935 SNPs, LDetect-like blocks U[400,1200], float32 LD"),
    "small": dict(model="viprs", n_blocks=16, block=4096, ld_dtype="int8", G=1, K=1, default_steps=300, fp="float32",
                  desc="65,536 SNPs = 16 LD blocks x 4096, int8 LD (CPU-baseline slice of c2)"),
    "ln": dict(model="viprs", n_blocks=0, block=0, ld_dtype="int8", G=1, K=1, M=1101824, lognormal=True, default_steps=300,
               fp="float32", desc="VIPRS spike-and-slab, 1,101,824 SNPs in LDetect-like log-normal LD blocks (median ~650 SNPs), "
                                  "int8 LD, float32, G=1"),
    "c5": dict(model="viprs", n_blocks=586, block=10240, ld_dtype="float64", G=1, K=1, default_steps=10, fp="float64",
               per_rank_blocks=73,
               desc="VIPRS float64, 10,240-SNP LD blocks (~5,120 nnz/row, float64 LD, tiled sweep); 586 blocks = 6.0 M SNPs over 8 "
                    "GPUs, 73 blocks (30 GB of LD) per GPU"),
}
ESIZE = {"int8": 1, "int16": 2, "float32": 4, "float64": 8}
FP32_FMA_PER_S = 148 * 128 * 1.965e9        # B200 CUDA-core FP32 FMA peak at the maximum SM clock
KERNEL_NAME = {"viprs": "vb::sweep_fast_kernel<int8,SlabModel>", "mix": "vb::sweep_fast_kernel<int16,MixModel<4>>",
               "grid": "vb::grid_sweep_kernel<float,int8>"}


def algorithmic_bytes(M, nnz, ld_dtype, tsize, G=1, K=1):
    """LD once + (indptr int64, left_bound int32) + (beta, n) + 7 state words per SNP x model
    (mixture: 2K + 5)."""
    state = (2 * K + 5) if K > 1 else 7 * G
    return nnz * ESIZE[ld_dtype] + M * 12 + 2 * M * tsize + M * state * tsize


def grid_hyper(M):
    """16 pi x 16 sigma_epsilon grid, sigma_epsilon-major (HyperparameterGrid.py:146-163,193-205,238-245)."""
    from scipy.stats import norm
    pis = np.logspace(np.log10(max(10. / M, 1e-5)), np.log10(min(1e4 / M, 0.2)), 16)
    p0 = max(0.1, norm.cdf((1e-5 - 0.1) / 0.1))
    ses = 1. - norm.ppf(np.linspace(p0, 0.9, 16), 0.1, 0.1)
    return [{"pi": float(p), "sigma_epsilon": float(s)} for s in ses for p in pis]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t0 is not None and not (t0 <= t <= t1):
                continue
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# workload geometry / sharding
# ---------------------------------------------------------------------------------------------------------------
def genome_sizes(wl, world=1):
    """Block sizes of the whole (fixed) genome of a workload."""
    from viprs_b200 import synth
    if wl.get("per_rank_blocks"):
        nb = min(wl["n_blocks"], wl["per_rank_blocks"] * world)
        return [wl["block"]] * nb
    if wl["n_blocks"]:
        return synth.block_sizes_for(wl["n_blocks"] * wl["block"], wl["block"])
    if wl.get("lognormal"):
        return synth.lognormal_sizes(wl["M"])
    return synth.ldetect_like_sizes(wl["M"])


def shard_blocks(sizes, rank, world):
    """Contiguous run of whole LD blocks of rank `rank`, balanced by sweep cost (viprs_b200.parallel)."""
    from viprs_b200 import parallel
    if world == 1:
        return 0, len(sizes)
    br = np.concatenate([[0], np.cumsum(sizes)])
    cut = parallel.partition_blocks(parallel.block_costs(br), world)
    return int(cut[rank]), int(cut[rank + 1])


def config_of(name, wl, scaling, world):
    """The workload as both arms (ours / reference) print it: static facts only, derived from the definition."""
    sizes = genome_sizes(wl, world)
    mult = world if (scaling == "weak" and not wl.get("per_rank_blocks")) else 1
    return {"workload": name + ": " + wl["desc"],
            "sharding": ("LD blocks of the fixed genome dealt to the ranks (strong scaling)" if scaling == "strong"
                         else "one genome-sized shard per rank (weak scaling)") if world > 1 else "single GPU",
            "snps": int(sum(sizes)) * mult, "ld_blocks": len(sizes) * mult,
            "nnz": int(sum(b * (b - 1) // 2 for b in sizes)) * mult,
            "grid_columns": wl["G"], "mixture_components": wl["K"],
            "step": "device-resident EM iteration: prepare + sweep + sums + scalar M-step/ELBO kernels"
                    + (" + one NCCL all-reduce" if world > 1 else "") + ", per-iteration scalars read back once per 64 steps"}


# ---------------------------------------------------------------------------------------------------------------
# CPU reference leg
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(name, wl, sweeps, warmup, threads, max_seconds=25.0):
    """The reference's own C++ sweep (oracle/_ref, compiled unmodified; the C port when it could not be compiled) on
    the host cores over the first 16 LD blocks of the workload -- the same arrays the GPU arm holds for those blocks."""
    import torch
    from oracle import cpu as ocpu
    from viprs_b200 import synth
    all_sizes = genome_sizes(wl)
    nb = min(16, len(all_sizes)) if name != "c5" else 2
    if name == "c1":
        nb = len(all_sizes)
    sizes = all_sizes[:nb]
    Mg = int(sum(all_sizes))
    T = np.float32 if wl["fp"] == "float32" else np.float64
    tdt = torch.float32 if wl["fp"] == "float32" else torch.float64
    inp = synth.make_inputs(sizes, ld_dtype=wl["ld_dtype"], float_dtype=tdt, device="cpu", M_total=Mg)
    M = int(sum(sizes))
    kind = "reference" if ocpu.have_ref() else "port"
    if kind == "port":
        threads = 1
    lb, ip, ld = inp["ld_left_bound"].numpy(), inp["ld_indptr"].numpy(), inp["ld_data"].numpy()
    beta, n = inp["std_beta"].numpy(), inp["n_per_snp"].numpy()
    G = 1
    if wl["model"] == "grid":
        G = 32                                    # a 32-column slice of the 256-column grid keeps the sample bounded
        recs = grid_hyper(Mg)[::8]
        pis, ses = np.array([r["pi"] for r in recs]), np.array([r["sigma_epsilon"] for r in recs])
        tau = pis * Mg / (1 - ses)
        vt = n[:, None] / ses + tau
        F = lambda a: np.asfortranarray(a.astype(T))
        ul, hv, mm = F(np.log(pis) - np.log1p(-pis) + .5 * (np.log(tau) - np.log(vt))), F(.5 * vt), F(n[:, None] / (vt * ses))
        st = {k: np.zeros((M, G), T, order="F") for k in ("var_mu", "eta", "q", "eta_diff")}
        st["var_gamma"] = F(np.tile(pis, (M, 1)))
        act = np.arange(G, dtype=np.int32)
        call = lambda: ocpu.e_step_grid(lb, ip, ld, beta, st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                        ul, hv, mm, inp["dq_scale"], act, threads, True, kind=kind)
        what = f"cpp e_step_grid<float,{wl['ld_dtype']}> on {G} of the 256 grid columns"
    elif wl["model"] == "mix":
        K = wl["K"]
        d = 2.0 ** np.linspace(-min(K - 1, 7), 0, K)
        pis, se = 0.01 * np.ones(K) / K, 0.8
        tau = d * (Mg * np.dot(1. / d, pis) / (1 - se))
        vt = n[:, None] / se + tau
        C = lambda a: np.ascontiguousarray(a.astype(T))
        ul, sv, mm = C(np.log(pis) - np.log1p(-pis) + .5 * (np.log(tau) - np.log(vt))), C(np.sqrt(.5 * vt)), C(n[:, None] / (vt * se))
        lnp = np.full(M, np.log(1 - pis.sum()), T)
        st = {"var_gamma": C(np.tile(pis, (M, 1))), "var_mu": np.zeros((M, K), T), "eta": np.zeros(M, T),
              "q": np.zeros(M, T), "eta_diff": np.zeros(M, T)}
        call = lambda: ocpu.e_step_mixture(lb, ip, ld, beta, st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                           lnp, ul, sv, mm, inp["dq_scale"], threads, True, kind=kind)
        what = f"cpp e_step_mixture<float,{wl['ld_dtype']}> K={K}"
    else:
        pi, se = 0.01, 0.8
        u_logs, shvt, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, pi * Mg / (1 - se), float_dtype=tdt)
        st = {k: np.zeros(M, T) for k in ("var_mu", "eta", "q", "eta_diff")}
        st["var_gamma"] = np.full(M, pi, T)
        call = lambda: ocpu.e_step(lb, ip, ld, beta, st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                   u_logs.numpy(), shvt.numpy(), mm.numpy(), inp["dq_scale"], threads, True, kind=kind)
        what = f"cpp e_step<{wl['fp']},{wl['ld_dtype']}>"
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + sweeps):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > max_seconds and len(times) >= 1:
            break
    sec = float(np.median(times))
    return {"value": M * G / sec, "unit": "SNP-updates/s", "cores": threads, "kind": kind,
            "sample": f"median of {len(times)} sweeps (after {warmup} warm-up) of {what}, low_memory=True, {M} SNPs = the first "
                      f"{len(sizes)} LD blocks of the workload's own arrays, {threads} OpenMP thread(s); {sec * 1e3:.1f} ms/sweep "
                      f"(min {min(times) * 1e3:.1f}, max {max(times) * 1e3:.1f})",
            "ms_per_sweep": sec * 1e3, "M": M}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def build_model(name, wl, rank, world, scaling):
    import torch
    from viprs_b200 import synth
    from viprs_b200.model import VIPRS, VIPRSGrid, VIPRSMix
    sizes = genome_sizes(wl, world)
    Mg = int(sum(sizes))
    if wl.get("per_rank_blocks"):
        # defined per GPU (c5: 30 GB of float64 LD each): rank r holds blocks [r p, (r + 1) p) of the genome
        b0 = min(rank * wl["per_rank_blocks"], len(sizes))
        b1 = min(b0 + wl["per_rank_blocks"], len(sizes))
        ids = list(range(b0, b1))
        mine = sizes[b0:b1]
    elif scaling == "strong":
        b0, b1 = shard_blocks(sizes, rank, world)
        ids = list(range(b0, b1))
        mine = sizes[b0:b1]
    else:
        ids = [rank * len(sizes) + i for i in range(len(sizes))]        # distinct blocks per rank
        mine = sizes
        Mg = Mg * world
    tdt = torch.float32 if wl["fp"] == "float32" else torch.float64
    inp = synth.make_inputs(mine, ld_dtype=wl["ld_dtype"], float_dtype=tdt, device="cuda", block_ids=ids, M_total=Mg)
    data = {1: dict(ld_data=inp["ld_data"], ld_indptr=inp["ld_indptr"], ld_left_bound=inp["ld_left_bound"],
                    std_beta=inp["std_beta"], n_per_snp=inp["n_per_snp"])}
    M = int(sum(mine))
    kw = dict(data=data, float_precision=wl["fp"], presharded=world > 1)
    if wl["model"] == "grid":
        m = VIPRSGrid(grid=grid_hyper(Mg), **kw)
    elif wl["model"] == "mix":
        m = VIPRSMix(K=wl["K"], **kw)
    else:
        m = VIPRS(**kw)
    del inp, data
    torch.cuda.empty_cache()
    return m, M, Mg, len(mine)


def profiled_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep launch from the newest ncu launch list committed under
    profiles/ for this workload (r<NN>*_<workload>_launches.csv); (None, None) when there is none."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r[0-9][0-9]*_{name}_launches*.csv")))
    for path in reversed(files):
        per_id = {}
        try:
            with open(path, newline="") as f:
                rows = [r for r in csv.reader(f) if len(r) >= 15 and r[0].isdigit()]
        except Exception:
            continue
        for r in rows:
            kname, metric, val = r[4], r[12], r[14]
            if "sweep" not in kname or "backward" in kname:
                continue
            if metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                per_id.setdefault(r[0], 0.0)
                per_id[r[0]] += float(val.replace(",", ""))
        if per_id:
            return float(np.mean(list(per_id.values()))), os.path.relpath(path, ROOT)
    return None, None


def run_workload(name, wl, steps, warmup, rank, world, local_rank, scaling, dist, want_e2e, use_graph=False):
    """Device-resident EM steps of one workload; returns the measurements (max over ranks)."""
    import torch
    tsize = 4 if wl["fp"] == "float32" else 8
    G, K = wl["G"], wl["K"]
    model, M, Mg, n_blocks = build_model(name, wl, rank, world, scaling)
    nnz = int(model.ld.nnz)
    if wl["model"] == "grid":
        model._batched = True
        model._init_grid_hyper({})
        model.initialize_variational_parameters()
        model._active = list(range(G))
    elif wl["model"] == "mix":
        model.initialize({"pis": list(0.01 * np.ones(K) / K), "sigma_epsilon": 0.8})
    else:
        model.initialize({"pi": 0.01, "sigma_epsilon": 0.8})

    # CUDA events around the sweep launch(es) inside every step: the roofline is for that kernel alone
    sweep_ev = []
    orig_sweep = model._sweep

    def timed_sweep():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); orig_sweep(); b.record()
        sweep_ev.append((a, b))
    model._sweep = timed_sweep
    n_ph = int(model.ld.n_phases)
    launches_per_step = 3 + (n_ph if n_ph == 1 else 2 * n_ph + 1)     # prepare + sums + em_update + sweep launches (tiled: + products)

    ld_bytes = nnz * ESIZE[wl["ld_dtype"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if ld_bytes < (256 << 20) else None

    # device-resident EM iterations (model.em_iterations): prepare -> sweep -> sums -> [all-reduce] -> scalar M-step / ELBO
    # kernels back to back, the host reads the per-iteration scalars once per chunk of <= 64 iterations
    if flush is not None:
        model._iter_hook = flush.zero_       # inputs smaller than the 126 MB L2: evict them between steps

    def run_steps(n):
        while n > 0:
            c = min(n, 64)
            model.em_iterations(c, graph=use_graph)
            n -= c

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    run_steps(warmup)
    barrier()
    sweep_ev.clear()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    run_steps(steps)
    e1.record()
    barrier()
    t_host1 = time.perf_counter()
    clocks = sampler.stop(t_host0, t_host1)
    if use_graph:
        # CUDA events cannot bracket a kernel inside a replayed graph: time the sweep launches of a few plain-launch
        # iterations right after the timed region instead (same kernels, same data)
        sweep_ev.clear()
        for _ in range(min(steps, 20)):
            model.em_iterations(1, graph=False)
        torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in sweep_ev]))
    stats = torch.tensor([total_ms, kern_ms], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(M), float(nnz), float(n_blocks)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, kern_ms = float(stats[0].item()), float(stats[1].item())
    M_all, nnz_all, nb_all = (int(v) for v in tot.tolist())
    model._sweep = orig_sweep
    out = {"value": M_all * G * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "kernel_ms": kern_ms,
           "steps": steps, "snps": M_all, "ld_blocks": nb_all, "nnz": nnz_all, "snps_this_rank": M, "clocks": clocks,
           "launches_per_step": launches_per_step, "flushed_l2": flush is not None,
           "kernel_ms_source": "CUDA events around every sweep launch inside the timed region" if not use_graph else
                               "CUDA events around the sweep launches of 20 plain-launch iterations right after the timed region "
                               "(the timed region replays a CUDA graph)"}
    # roofline of the dominant kernel (the sweep): whole-job algorithmic bytes over the slowest rank's launch time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    abytes = algorithmic_bytes(M_all, nnz_all, wl["ld_dtype"], tsize, G, K)
    achieved = abytes / (kern_ms * 1e-3) / 1e9 / world           # per GPU
    traffic, tsrc = profiled_traffic(name)
    kname = KERNEL_NAME[wl["model"]] if name != "c5" else "vb::sweep_kernel<double,double,SlabModel> (+ backward_dot / forward_axpy products between tiles)"
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "kernel": kname, "kernel_ms": kern_ms,
                       "algorithmic_bytes_per_launch": abytes / world,
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                       "traffic_source": tsrc}
    if wl["model"] == "grid":
        out["roofline"]["fp32_pipe_frac"] = 2.0 * nnz_all * G / (kern_ms * 1e-3) / FP32_FMA_PER_S / world
        out["roofline"]["note"] = ("c3 is bound by the CUDA-core FP32 pipe (2*nnz*G FMA per sweep; tensor cores excluded by the "
                                   "north star), not by HBM: fp32_pipe_frac is the binding fraction")

    # ---- e2e: ONE C-ABI call per step with HOST (pinned) state arrays on the resident LD ----
    if want_e2e and wl["model"] in ("viprs", "mix") and M > 0:
        from viprs_b200 import e_step as es
        model._materialize_q()
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype).pin_memory().copy_(t)
        names_in = ["std_beta_dev", "_g", "_mu", "_eta", "_q", "_ul", "_tt", "_mm"] + (["_lnp"] if wl["model"] == "mix" else [])
        host = {k: pin(getattr(model, k)) for k in names_in}
        host["_diff"] = pin(model._diff)
        dq = model.dequantize_scale

        def e2e_step():
            if wl["model"] == "mix":
                es.cpp_e_step_mixture_resident(model.ld, host["std_beta_dev"], host["_g"], host["_mu"], host["_eta"], host["_q"],
                                               host["_diff"], host["_lnp"], host["_ul"], host["_tt"], host["_mm"], dq, False)
            else:
                es.cpp_e_step_resident(model.ld, host["std_beta_dev"], host["_g"], host["_mu"], host["_eta"], host["_q"],
                                       host["_diff"], host["_ul"], host["_tt"], host["_mm"], dq, False)

        e_steps = min(steps, 100)
        for _ in range(3):
            e2e_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e_steps):
            e2e_step()
        t1.record()
        barrier()
        ems = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems = float(ems.item())
        h2d = sum(host[k].numel() * host[k].element_size() for k in names_in)
        d2h = sum(host[k].numel() * host[k].element_size() for k in ("_g", "_mu", "_eta", "_q", "_diff"))
        out["e2e"] = {"value": M_all * G * e_steps / (ems * 1e-3), "unit": "SNP-updates/s",
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / e_steps,
                      "steps": e_steps,
                      "call": "viprs_b200_cpp_e_step%s_resident (host state arrays of this rank's shard; q in/out with the "
                              "reference's incremental bookkeeping, nothing vouched: sweep + update_q_factor pass, in row chunks "
                              "on internal streams so that copies and sweeps overlap)" % ("_mixture" if wl["model"] == "mix" else "")}
    del model
    torch.cuda.empty_cache()
    return out


def multi_gpu_parity(rank, world):
    """N > 1 only (the driver's GPU test box has one GPU, so tests/test_multi_gpu.py is skipped there): 8 EM iterations of
    VIPRS on a small fixed genome (two chromosomes, ragged LD blocks), sharded over the ranks of this run with one
    all-reduce per iteration, against the same iterations on rank 0 alone.  Outside every timed region."""
    import torch
    from viprs_b200 import synth
    from viprs_b200.model import VIPRS
    sizes = [700, 512, 900, 333, 1200, 64, 800, 1024, 256, 2048, 96, 640, 1500, 300, 450, 777]
    inp = synth.make_inputs(sizes, ld_dtype="int8", float_dtype=torch.float32, device="cpu", seed=11, n=50000)
    half = sum(sizes[:8])
    ip, lb = inp["ld_indptr"].numpy(), inp["ld_left_bound"].numpy()

    def chrom(r0, r1):
        return dict(ld_data=inp["ld_data"].numpy()[ip[r0]:ip[r1]], ld_indptr=(ip[r0:r1 + 1] - ip[r0]),
                    ld_left_bound=(lb[r0:r1] - r0).astype(np.int32), std_beta=inp["std_beta"].numpy()[r0:r1],
                    n_per_snp=inp["n_per_snp"].numpy()[r0:r1])

    data = {1: chrom(0, half), 2: chrom(half, sum(sizes))}

    def history(shard, theta_0):
        # theta_0 = None: the reference's random initialisation (VIPRS.py:260-265), drawn by every rank from a DIFFERENT
        # numpy seed in the sharded run: rank 0's draw must reach all shards (broadcast in initialize_theta)
        np.random.seed(4321 + (rank if shard else 0))
        m = VIPRS(data=data, float_precision="float32", shard=shard)
        m.initialize(theta_0)
        h = []
        for _ in range(8):
            m.e_step(); m.m_step()
            h.append([m.elbo(), float(m.pi), float(m.sigma_epsilon), float(m.tau_beta), m.max_eta_diff()])
        return np.array(h)

    fixed = {"pi": 0.02, "sigma_epsilon": 0.8}
    sharded, sharded_rand = history(True, dict(fixed)), history(True, None)
    if rank != 0:
        return None
    alone, alone_rand = history(False, dict(fixed)), history(False, None)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-30)))
    err, err_rand = rel(alone, sharded), rel(alone_rand, sharded_rand)
    return {"what": "8 EM iterations (ELBO, pi, sigma_epsilon, tau_beta, max|eta_diff|) of VIPRS on a %d-SNP genome sharded "
                    "over %d GPUs vs the same on one GPU; once with fixed initial hyper-parameters, once with the random "
                    "initialisation drawn from a different seed on every rank" % (sum(sizes), world),
            "max_rel_diff": err, "max_rel_diff_random_init": err_rand, "tolerance": 1e-12,
            "ok": bool(err <= 1e-12 and err_rand <= 1e-12)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the EM iteration as a CUDA graph in the timed region")
    ap.add_argument("--no-extras", action="store_true", help="skip the c3 / c4 / c1 sub-measurements of the default run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    steps = args.steps if args.steps > 0 else wl["default_steps"]
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    scaling = args.scaling if args.workload != "c5" else "weak"       # c5 is defined per GPU (30 GB of LD each)
    dtype = "f32" if wl["fp"] == "float32" else "f64"

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        base = cpu_reference_leg(args.workload, wl, max(3, min(steps, 15)), min(max(args.warmup, 1), 2), threads, max_seconds=120.0)
        line = {"impl": "reference", "metric": "E-step SNP-updates/s", "value": base["value"], "unit": "SNP-updates/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": base["ms_per_sweep"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": config_of(args.workload, wl, scaling, max(args.gpus, 1)),
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "SNP-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import viprs_b200  # noqa: F401  (fails loudly when the CUDA extension is missing)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: viprs_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # (NCCL_DEBUG is left as the caller set it: with VERSION or above NCCL prints its banner on stdout next to rank 0's
        # ONE JSON line, which is the last line)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    main_res = run_workload(args.workload, wl, steps, warmup, rank, world, local_rank, scaling, dist, not args.no_e2e,
                            use_graph=args.graph)

    extras = {}
    if args.workload == "c2" and not args.no_extras:
        for nm in ("c3", "c4", "c1"):
            w = WORKLOADS[nm]
            try:
                r = run_workload(nm, w, w["default_steps"] if nm != "c4" else 50, 3, rank, world, local_rank, scaling, dist, False,
                                 use_graph=args.graph)
                extras[nm] = {"workload": nm + ": " + w["desc"], "value": r["value"], "unit": "SNP-updates/s",
                              "ms_per_step": r["ms_per_step"], "kernel_ms": r["kernel_ms"], "steps": r["steps"],
                              "frac": r["roofline"]["frac"], "fp32_pipe_frac": r["roofline"].get("fp32_pipe_frac"),
                              "kernel": r["roofline"]["kernel"], "snps": r["snps"], "flushed_l2": r["flushed_l2"]}
            except Exception as ex:          # an extra must never take the headline down with it
                extras[nm] = {"error": repr(ex)[:300]}

    mgp = None
    if world > 1 and not args.no_extras:
        try:
            mgp = multi_gpu_parity(rank, world)
        except Exception as ex:
            mgp = {"error": repr(ex)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_leg(args.workload, wl, 10, 1, os.cpu_count() or 1, max_seconds=20.0)
        cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        b1 = cpu_reference_leg(args.workload, wl, 5, 1, 1, max_seconds=15.0)
        cpu["threads_1"] = {k: b1[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        cfg = config_of(args.workload, wl, scaling, world)
        assert (cfg["snps"], cfg["ld_blocks"], cfg["nnz"]) == (main_res["snps"], main_res["ld_blocks"], main_res["nnz"]), \
            (cfg, main_res["snps"], main_res["ld_blocks"], main_res["nnz"])
        run_info = {"snps_rank0": main_res["snps_this_rank"],
                    "l2": ("inputs (%.2f GB of LD per sweep and GPU) are larger than the 126 MB L2; no flush needed"
                           % (main_res["nnz"] * ESIZE[wl["ld_dtype"]] / world / 1e9))
                          if not main_res["flushed_l2"] else "L2 flushed between steps by writing a 256 MiB buffer (inside the timed region)"}
        line = {"metric": "E-step SNP-updates/s", "value": main_res["value"], "unit": "SNP-updates/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": main_res["ms_per_step"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": cfg, "roofline": main_res["roofline"], "cpu_baseline": cpu, "e2e": main_res.get("e2e"),
                "gpu_launches": main_res["launches_per_step"] * steps, "clocks": main_res["clocks"], "run": run_info}
        line["roofline"]["kernel_ms_source"] = main_res["kernel_ms_source"]
        if extras:
            line["workloads"] = extras
        if mgp is not None:
            line["multi_gpu_parity"] = mgp
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
