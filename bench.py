#!/usr/bin/env python
"""
bench.py -- E-step SNP-updates/s of the B200 coordinate-ascent sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c1|small]

One "step" = the E-step of one EM iteration over the whole per-rank genome, exactly as `fit()` runs it: the per-SNP
pre-compute (prepare kernel), the Gauss-Seidel sweep over every LD block, the M-step / ELBO reductions (sums kernel),
the read-back of those few hundred bytes and the scalar M-step on the host; with N > 1 also the one small NCCL
all-reduce.  Workloads (synthetic block-diagonal PD LD, seed 7209):
  c2 (default, BASELINE.json configs[1]): VIPRS spike-and-slab, 1,101,824 SNPs = 269 LD blocks x 4096, int8 LD
      (2047.5 stored entries per row, upper-triangular), float32 state, G = 1.
  c3: VIPRSGrid, same LD, 256 (pi x sigma_epsilon) grid columns sharing every LD row (e_step_grid semantics).
  c4: VIPRSMix K = 4, int16 LD.       c1: chr22-shaped 15,935 SNPs, float32 LD.       small: 16 blocks of c2.
Multi-GPU is weak scaling: every rank owns its own genome-sized shard of whole LD blocks (N x 1.1M SNPs in total);
the hyper-parameters are global (one all-reduce per step).

Keys of the JSON line (rank 0):
  value      SNP-updates/s (SNPs x grid columns x steps / device time, max over ranks), everything resident in HBM.
  roofline   the sweep kernel alone: algorithmic bytes per launch / mean launch duration (CUDA events
             around the launch, inside the timed region) / measured HBM copy peak (MEASURED_PEAKS.json).  For c3 the
             binding resource is the CUDA-core FP32 pipe, reported as fp32_pipe_frac next to the HBM fraction.
  e2e        the same step through the reference-facing call with HOST (pinned) buffers: what VIPRS.e_step() hands to
             cpp_e_step every iteration goes host->device, what m_step()/elbo() read goes device->host, all inside the
             timed region.
  cpu_baseline  the reference's own C++ e_step (compiled unmodified, oracle/_ref) on the host cores, 16-block sample.
"""
import argparse
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
    "c2": dict(model="viprs", n_blocks=269, block=4096, ld_dtype="int8", G=1, K=1, default_steps=300,
               desc="VIPRS spike-and-slab, 1,101,824 SNPs = 269 LD blocks x 4096, int8 LD (2047.5 nnz/row), float32, G=1"),
    "c3": dict(model="grid", n_blocks=269, block=4096, ld_dtype="int8", G=256, K=1, default_steps=5,
               desc="VIPRSGrid 256 (16 pi x 16 sigma_epsilon) columns sharing LD rows, 1,101,824 SNPs = 269 LD blocks x 4096, "
                    "int8 LD, float32"),
    "c4": dict(model="mix", n_blocks=269, block=4096, ld_dtype="int16", G=1, K=4, default_steps=100,
               desc="VIPRSMix K=4, 1,101,824 SNPs = 269 LD blocks x 4096, int16 LD, float32"),
    "c1": dict(model="viprs", n_blocks=0, block=0, ld_dtype="float32", G=1, K=1, M=15935, default_steps=300,
               desc="VIPRS spike-and-slab, chr22-shaped 15,935 SNPs, LDetect-like blocks U[400,1200], float32 LD"),
    "small": dict(model="viprs", n_blocks=16, block=4096, ld_dtype="int8", G=1, K=1, default_steps=300,
                  desc="65,536 SNPs = 16 LD blocks x 4096, int8 LD (CPU-baseline slice of c2)"),
}
ESIZE = {"int8": 1, "int16": 2, "float32": 4, "float64": 8}
FP32_FMA_PER_S = 148 * 128 * 1.965e9        # B200 CUDA-core FP32 FMA peak at the maximum SM clock


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


def make_sizes(wl, nb=None):
    from viprs_b200 import synth
    if wl["n_blocks"]:
        return synth.block_sizes_for((nb or wl["n_blocks"]) * wl["block"], wl["block"])
    return synth.ldetect_like_sizes(wl["M"])


def cpu_reference_leg(wl, steps, warmup, threads, max_seconds=25.0):
    """The reference's own C++ sweep (oracle/_ref, compiled unmodified; the C port when it could not be compiled) on
    the host cores over a 16-block slice of the workload."""
    import torch
    from oracle import cpu as ocpu
    from viprs_b200 import synth
    sizes = make_sizes(wl, min(16, wl["n_blocks"]) if wl["n_blocks"] else None)
    inp = synth.make_inputs(sizes, ld_dtype=wl["ld_dtype"], float_dtype=torch.float32, device="cpu")
    M = int(sum(sizes))
    kind = "reference" if ocpu.have_ref() else "port"
    if kind == "port":
        threads = 1
    lb, ip, ld = inp["ld_left_bound"].numpy(), inp["ld_indptr"].numpy(), inp["ld_data"].numpy()
    beta, n = inp["std_beta"].numpy(), inp["n_per_snp"].numpy()
    T = np.float32
    G = 1
    if wl["model"] == "grid":
        G = 32                                    # a 32-column slice of the 256-column grid keeps the sample bounded
        recs = grid_hyper(wl["n_blocks"] * wl["block"])[::8]
        pis, ses = np.array([r["pi"] for r in recs]), np.array([r["sigma_epsilon"] for r in recs])
        tau = pis * M / (1 - ses)
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
        tau = d * (M * np.dot(1. / d, pis) / (1 - se))
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
        u_logs, shvt, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, pi * M / (1 - se))
        st = {k: np.zeros(M, T) for k in ("var_mu", "eta", "q", "eta_diff")}
        st["var_gamma"] = np.full(M, pi, T)
        call = lambda: ocpu.e_step(lb, ip, ld, beta, st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                   u_logs.numpy(), shvt.numpy(), mm.numpy(), inp["dq_scale"], threads, True, kind=kind)
        what = f"cpp e_step<float,{wl['ld_dtype']}>"
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > max_seconds and len(times) >= 1:
            break
    sec = float(np.mean(times))
    return {"value": M * G / sec, "unit": "SNP-updates/s", "cores": threads, "kind": kind,
            "sample": f"{len(times)} sweeps (after {warmup} warm-up) of {what}, low_memory=True, {M} SNPs = {len(sizes)} LD "
                      f"blocks of the same synthetic workload, {threads} OpenMP thread(s); {sec * 1e3:.1f} ms/sweep",
            "ms_per_sweep": sec * 1e3, "M": M}


def build_model(wl, rank, world):
    import torch
    from viprs_b200 import synth
    from viprs_b200.model import VIPRS, VIPRSGrid, VIPRSMix
    sizes = make_sizes(wl)
    inp = synth.make_inputs(sizes, ld_dtype=wl["ld_dtype"], float_dtype=torch.float32, device="cuda", seed=synth.SEED + rank)
    data = {1: dict(ld_data=inp["ld_data"], ld_indptr=inp["ld_indptr"], ld_left_bound=inp["ld_left_bound"],
                    std_beta=inp["std_beta"], n_per_snp=inp["n_per_snp"])}
    M = int(sum(sizes))
    kw = dict(data=data, float_precision="float32", presharded=world > 1)
    if wl["model"] == "grid":
        m = VIPRSGrid(grid=grid_hyper(M * world), **kw)
    elif wl["model"] == "mix":
        m = VIPRSMix(K=wl["K"], **kw)
    else:
        m = VIPRS(**kw)
    del inp, data
    torch.cuda.empty_cache()
    return m, M, len(sizes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    steps = args.steps if args.steps > 0 else wl["default_steps"]
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tsize = 4
    G, K = wl["G"], wl["K"]

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        base = cpu_reference_leg(wl, max(1, min(steps, 3)), min(args.warmup, 1), threads, max_seconds=120.0)
        line = {"impl": "reference", "metric": "E-step SNP-updates/s", "value": base["value"], "unit": "SNP-updates/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": base["ms_per_sweep"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"]},
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
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    model, M, n_blocks = build_model(wl, rank, world)
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

    # CUDA events around the sweep launch inside every step: the roofline is for that kernel alone
    sweep_ev = []
    orig_sweep = model._sweep

    def timed_sweep():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); orig_sweep(); b.record()
        sweep_ev.append((a, b))
    model._sweep = timed_sweep
    launches_per_step = 3          # prepare + sweep + sums (the all-reduce is NCCL's kernel, not ours)

    ld_bytes = nnz * ESIZE[wl["ld_dtype"]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if ld_bytes < (256 << 20) else None

    def step():
        if flush is not None:
            flush.zero_()          # inputs smaller than the 126 MB L2: evict them between steps
        model.e_step()
        model.m_step()             # sums kernel, (all-reduce), read-back of the reduced table, scalar M-step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sweep_ev.clear()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    t_host1 = time.perf_counter()
    clocks = sampler.stop(t_host0, t_host1)
    total_ms = e0.elapsed_time(e1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in sweep_ev]))
    if world > 1:
        t = torch.tensor([total_ms, kern_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms / steps
    value = M * G * world * steps / (total_ms * 1e-3)
    model._sweep = orig_sweep

    # ---- roofline of the dominant kernel (the sweep) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    abytes = algorithmic_bytes(M, nnz, wl["ld_dtype"], tsize, G, K)
    achieved = abytes / (kern_ms * 1e-3) / 1e9
    kname = {"viprs": "vb::sweep_fast_kernel<int8,SlabModel>", "mix": "vb::sweep_fast_kernel<int16,MixModel<4>>",
             "grid": "vb::grid_sweep_kernel<float,int8>"}[wl["model"]]
    # dram__bytes_read.sum + dram__bytes_write.sum per sweep launch from the ncu passes committed under profiles/
    traffic = {"c2": 2.337e9, "c3": 16.82e9, "c4": 4.668e9}.get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kname, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": abytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "traffic_source": {"c2": "profiles/r01b_c2_launches.csv", "c3": "profiles/r01_c3_launches_dram.csv",
                                   "c4": "profiles/r01b_c4_launches.csv"}.get(args.workload)}
    if wl["model"] == "grid":
        fma = 2.0 * nnz * G
        roofline["fp32_pipe_frac"] = fma / (kern_ms * 1e-3) / FP32_FMA_PER_S
        roofline["note"] = ("c3 is bound by the CUDA-core FP32 pipe (2*nnz*G FMA per sweep; tensor cores excluded by the "
                            "north star), not by HBM: fp32_pipe_frac is the binding fraction")

    # ---- e2e: the reference-facing call with HOST (pinned) buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h2d_t = [model._ul, model._tt, model._mm]                   # what VIPRS.e_step() recomputes on the host per iteration
        d2h_t = [model._g, model._mu, model._eta, model._diff]      # what m_step() / elbo() read back
        hb_in = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in h2d_t]
        hb_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in d2h_t]
        for h, t in zip(hb_in, h2d_t):
            h.copy_(t)
        e_steps = steps if wl["model"] != "grid" else min(steps, 5)

        def e2e_step():
            for h, t in zip(hb_in, h2d_t):
                t.copy_(h, non_blocking=True)
            orig_sweep()
            for h, t in zip(hb_out, d2h_t):
                h.copy_(t, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e_steps):
            e2e_step()
        t1.record()
        barrier()
        ems = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": M * G * world * e_steps / (ems * 1e-3), "unit": "SNP-updates/s",
               "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in h2d_t)),
               "d2h_bytes_per_step": int(sum(t.numel() * t.element_size() for t in d2h_t)),
               "ms_per_step": ems / e_steps, "steps": e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_leg(wl, 3, 1, os.cpu_count() or 1)
        cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "E-step SNP-updates/s", "value": value, "unit": "SNP-updates/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"], "snps_per_gpu": M, "ld_blocks_per_gpu": n_blocks,
                           "nnz_per_gpu": nnz, "grid_columns": G, "mixture_components": K,
                           "step": "prepare + sweep + sums kernels, read-back of the reduced sums, scalar M-step"
                                   + (", one NCCL all-reduce" if world > 1 else ""),
                           "l2": ("inputs (%.2f GB of LD per sweep) are larger than the 126 MB L2; no flush needed" % (ld_bytes / 1e9))
                                 if flush is None else "L2 flushed between steps by writing a 256 MiB buffer (inside the timed region)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * steps, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
