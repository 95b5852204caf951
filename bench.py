#!/usr/bin/env python
"""
bench.py -- E-step SNP-updates/s of the B200 coordinate-ascent sweep (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|...]

One "step" = one EM iteration's E-step over the whole (per-rank) genome: the per-SNP pre-compute, the
Gauss-Seidel sweep over every LD block, and the M-step/ELBO reductions (one tiny NCCL all-reduce when
N > 1).  Workload c2 (default) is BASELINE.json configs[1]: VIPRS spike-and-slab, 1,101,824 SNPs
(269 LD blocks x 4096), int8 LD (~2k stored entries per row, upper-triangular), float32 state, G = 1.
Multi-GPU is weak scaling: every rank owns its own 269-block shard (N x 1.1M SNPs genome-wide).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
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
    # name: (n_blocks, block, ld_dtype, float)  -- SURVEY.md section 8(d)
    "c2": dict(n_blocks=269, block=4096, ld_dtype="int8", fdt="float32",
               desc="VIPRS spike-and-slab, 1,101,824 SNPs = 269 LD blocks x 4096, int8 LD (2047.5 nnz/row), float32, G=1"),
    "c1": dict(n_blocks=0, block=0, ld_dtype="float32", fdt="float32", M=15935,
               desc="VIPRS spike-and-slab, chr22-shaped 15,935 SNPs, LDetect-like blocks U[400,1200], float32 LD"),
    "small": dict(n_blocks=16, block=4096, ld_dtype="int8", fdt="float32",
                  desc="65,536 SNPs = 16 LD blocks x 4096, int8 LD (CPU-baseline slice of c2)"),
}
ESIZE = {"int8": 1, "int16": 2, "float32": 4, "float64": 8}


def algorithmic_bytes(M, nnz, ld_dtype, tsize, G=1):
    """LD once + (indptr int64, left_bound int32) + (beta, n) + 7 state words."""
    return nnz * ESIZE[ld_dtype] + M * 12 + 2 * M * tsize + M * G * 7 * tsize


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(wl, steps, warmup, threads, max_seconds=25.0):
    """The reference's own C++ e_step (oracle/_ref, compiled unmodified) -- or the C port when the
    reference could not be compiled -- timed on the host cores over a 16-block slice of the workload."""
    import torch
    from oracle import cpu as ocpu
    from viprs_b200 import synth
    nb = min(16, wl["n_blocks"]) if wl["n_blocks"] else 0
    sizes = synth.block_sizes_for(nb * wl["block"], wl["block"]) if nb else synth.ldetect_like_sizes(wl["M"])
    inp = synth.make_inputs(sizes, ld_dtype=wl["ld_dtype"], float_dtype=torch.float32, device="cpu")
    M = int(sum(sizes))
    pi, se = 0.01, 0.8
    tau = pi * M / (1 - se)
    u_logs, shvt, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, tau)
    st = {k: np.zeros(M, np.float32) for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = np.full(M, pi, np.float32)
    kind = "reference" if ocpu.have_ref() else "port"
    if kind == "port":
        threads = 1
    args = (inp["ld_left_bound"].numpy(), inp["ld_indptr"].numpy(), inp["ld_data"].numpy(), inp["std_beta"].numpy(),
            st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"], u_logs.numpy(), shvt.numpy(), mm.numpy(),
            inp["dq_scale"], threads, True)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ocpu.e_step(*args, kind=kind)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > max_seconds and len(times) >= 1:
            break
    sec = float(np.mean(times))
    return {"value": M / sec, "unit": "SNP-updates/s", "cores": threads, "kind": kind,
            "sample": f"{len(times)} sweeps (after {min(warmup, 1)}+ warm-up) of cpp e_step<float,{wl['ld_dtype']}>, "
                      f"low_memory=True, {M} SNPs = {len(sizes)} LD blocks of the same synthetic workload, "
                      f"{threads} OpenMP thread(s); {sec * 1e3:.1f} ms/sweep",
            "ms_per_sweep": sec * 1e3, "M": M}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tsize = 4

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        base = cpu_reference_leg(wl, max(1, min(args.steps, 3)), min(args.warmup, 1), threads, max_seconds=120.0)
        line = {"impl": "reference", "metric": "E-step SNP-updates/s", "value": base["value"], "unit": "SNP-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_sweep"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"]},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "SNP-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import viprs_b200
    from viprs_b200 import synth
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: viprs_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic workload, generated directly in HBM (per-rank shard; seeds differ per rank) ----
    if wl["n_blocks"]:
        sizes = synth.block_sizes_for(wl["n_blocks"] * wl["block"], wl["block"])
    else:
        sizes = synth.ldetect_like_sizes(wl["M"])
    inp = synth.make_inputs(sizes, ld_dtype=wl["ld_dtype"], float_dtype=torch.float32, device="cuda",
                            seed=synth.SEED + rank)
    M = int(sum(sizes))
    ld = viprs_b200.DeviceLD(inp["ld_data"], inp["ld_indptr"], inp["ld_left_bound"])
    nnz = int(ld.nnz)
    del inp["ld_data"]
    torch.cuda.empty_cache()
    pi, se = 0.01, 0.8
    tau = pi * (M * world) / (1 - se)
    u_logs, shvt, mm, _ = synth.e_step_inputs(inp["std_beta"], inp["n_per_snp"], pi, se, tau)
    st = {k: torch.zeros(M, dtype=torch.float32, device="cuda") for k in ("var_mu", "eta", "q", "eta_diff")}
    st["var_gamma"] = torch.full((M,), pi, dtype=torch.float32, device="cuda")

    launches = [0]

    def step():
        viprs_b200.e_step_device(ld, inp["std_beta"], st["var_gamma"], st["var_mu"], st["eta"], st["q"], st["eta_diff"],
                                 u_logs, shvt, mm, inp["dq_scale"], False)
        launches[0] += 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    launches[0] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = M * world * args.steps / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (the sweep): algorithmic bytes / mean launch duration ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    abytes = algorithmic_bytes(M, nnz, wl["ld_dtype"], tsize)
    kern_ms = float(np.mean(per_launch_ms))
    achieved = abytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "vb::sweep_kernel", "algorithmic_bytes_per_launch": abytes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"}

    # ---- e2e: the reference-facing call with HOST (pinned) buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hb = {k: torch.empty(M, dtype=torch.float32).pin_memory() for k in
              ("std_beta", "var_gamma", "var_mu", "eta", "q", "eta_diff", "u_logs", "shvt", "mm")}
        hb["std_beta"].copy_(inp["std_beta"]); hb["u_logs"].copy_(u_logs); hb["shvt"].copy_(shvt); hb["mm"].copy_(mm)
        for k in ("var_gamma", "var_mu", "eta", "q", "eta_diff"):
            hb[k].copy_(st[k])
        h2d = ("u_logs", "shvt", "mm")                     # what VIPRS.e_step() recomputes on the host each iteration
        d2h = ("var_gamma", "var_mu", "eta", "eta_diff")   # what m_step()/elbo() read back
        dv = {"u_logs": u_logs, "shvt": shvt, "mm": mm}

        def e2e_step():
            for k in h2d:
                dv[k].copy_(hb[k], non_blocking=True)
            step()
            for k in d2h:
                hb[k].copy_(st[k], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            e2e_step()
        t1.record()
        barrier()
        ems = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": M * world * args.steps / (ems * 1e-3), "unit": "SNP-updates/s",
               "h2d_bytes_per_step": len(h2d) * M * 4, "d2h_bytes_per_step": len(d2h) * M * 4,
               "ms_per_step": ems / args.steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = cpu_reference_leg(wl, 3, 1, os.cpu_count() or 1)
        cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "E-step SNP-updates/s", "value": value, "unit": "SNP-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"], "snps_per_gpu": M, "ld_blocks_per_gpu": len(sizes),
                           "nnz_per_gpu": nnz, "grid_columns": 1,
                           "l2": "inputs (%.2f GB of LD per sweep) are larger than the 126 MB L2; no flush needed" % (nnz * ESIZE[wl["ld_dtype"]] / 1e9)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0], "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
