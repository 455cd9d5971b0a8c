#!/usr/bin/env python
"""Benchmark of the SR sampling hot path (BASELINE.json metric: events/s of SR sampling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload single_e|multipart] [--events B] [--n-steps S] [--precision bf16|fp32]

One "step" = one ``generate_samples`` pass (Euler, ``n_steps`` grid points = ``n_steps-1``
network evaluations) over one batch of ``B`` synthetic events per GPU.  Default workload =
BASELINE.json configs[1]: single-electron shapes, B = 4096 per GPU, n_steps = 25.

* ``value``  : whole-job events/s with packed inputs resident in HBM, timed with CUDA events
               around K calls of the C-ABI ``srhep_sample`` (max over ranks).
* ``e2e``    : the same metric through the public API (``FlowModel.generate_samples``) with
               the batch in pinned HOST memory: H2D of the batch, packing, binding, sampling,
               D2H of the final cells all inside the timed region (plus, for N > 1, the final
               NCCL gather of the outputs to rank 0).
* ``roofline``: dominant kernel category of one evaluation, timed live with CUDA events on
               the launching stream (srhep_profile), against MEASURED_PEAKS.json.
* ``cpu_baseline``: the CPU oracle port (oracle/sr_oracle.py, a restatement of the reference's
               PyTorch modules) on a bounded sample of the same workload, all host threads.
* ``--impl reference``: the reference arm = that CPU port timed as the whole run.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from superresolutionhep_b200.default_configs import flow_config          # noqa: E402
from superresolutionhep_b200.synthetic import (synthetic_events, synthetic_noise,  # noqa: E402
                                               synthetic_state_dict)

METRIC = "events/sec SR sampling (N ODE steps)"
UNIT = "events/s"
WEIGHT_SEED = 7


def flops_per_eval(counts: np.ndarray) -> float:
    """F(n) = 5 088 448 n + 6 144 n^2 + 3 215 360 per event per evaluation."""
    n = counts.astype(np.float64)
    return float((5088448.0 * n + 6144.0 * n * n + 3215360.0).sum())


def load_traffic(workload, events, precision):
    """DRAM bytes per launch of each kernel category from the committed ncu capture (profiles/), if it was taken
    on this exact workload; None otherwise."""
    p = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if not os.path.isfile(p):
        return {}
    with open(p) as fp:
        d = json.load(fp)
    if d.get("workload") != workload or d.get("events_per_gpu") != events or d.get("precision") != precision:
        return {}
    return d.get("dram_bytes_per_launch", {})


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fp:
            d = json.load(fp)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            try:
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": float(np.median(pw)) if pw else None}


def oracle_events_per_s(kind: str, n_events: int, n_steps: int, repeats: int = 1):
    """The CPU port on a bounded sample: ``n_events`` events, one padded batch, Euler."""
    from oracle import sr_oracle
    cfg = flow_config(kind)
    from superresolutionhep_b200.config import SrDims
    sd = synthetic_state_dict(SrDims.from_config(cfg), seed=WEIGHT_SEED)
    dims = sr_oracle.derive_dims(cfg)
    batch = synthetic_events(kind, n_events, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    torch.set_num_threads(os.cpu_count() or 1)
    best = None
    with torch.no_grad():
        sr_oracle.flow_forward(sd, dims, batch, x0, torch.zeros(n_events))            # warm-up evaluation
        for _ in range(repeats):
            t0 = time.perf_counter()
            sr_oracle.generate_samples(sd, dims, batch, x0, n_steps=n_steps, method="euler")
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_events / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ev = args.ref_events
    times = []
    from oracle import sr_oracle
    from superresolutionhep_b200.config import SrDims
    cfg = flow_config(args.workload)
    sd = synthetic_state_dict(SrDims.from_config(cfg), seed=WEIGHT_SEED)
    dims = sr_oracle.derive_dims(cfg)
    batch = synthetic_events(args.workload, n_ev, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            sr_oracle.generate_samples(sd, dims, batch, x0, n_steps=args.n_steps, method="euler")
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    val = n_ev * len(times) / total
    sample = (f"{n_ev} synthetic {args.workload} events per step (one padded batch), euler n_steps={args.n_steps} "
              f"({args.n_steps - 1} evaluations), fp32, torch CPU {torch.__version__}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, args.events, args.precision),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(args, events_per_gpu, precision):
    return {"workload": f"{args.workload} SR sampling, {events_per_gpu} synthetic events per GPU, euler n_steps={args.n_steps} "
                        f"({args.n_steps - 1} network evaluations per event)",
            "events_per_gpu": events_per_gpu, "n_steps": args.n_steps, "method": "euler", "precision": precision,
            "l2": "per-step activation traffic (GBs) and the packed state exceed the 126 MB L2; no explicit flush"}


def main():
    # Libraries (NCCL with NCCL_DEBUG=VERSION, torchrun banners) may write to stdout; the contract is ONE JSON line there.
    # Everything before the final print goes to stderr.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE is not None:
        print(_RESULT_LINE, flush=True)


_RESULT_LINE = None


def _emit(line: dict) -> None:
    global _RESULT_LINE
    _RESULT_LINE = json.dumps(line)


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="single_e", choices=["single_e", "multipart"])
    ap.add_argument("--events", type=int, default=4096, help="events per GPU per step")
    ap.add_argument("--n-steps", type=int, default=25)
    ap.add_argument("--precision", default=os.environ.get("SRHEP_BENCH_PRECISION", "bf16"), choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--ref-events", type=int, default=64, help="events per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--pass-tokens", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        run_reference(args)
        return
    if args.steps is None:
        args.steps = 3 if args.precision == "fp32" else 5

    import torch.distributed as dist
    from superresolutionhep_b200 import FlowModel, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.events
    cfg = flow_config(args.workload)
    model = FlowModel(cfg, precision=args.precision)
    model.load_state_dict(synthetic_state_dict(model.dims, seed=WEIGHT_SEED))
    model.eval().cuda(dev)
    model.pass_tokens = args.pass_tokens
    # every rank gets its own entry range of the synthetic sample (weak scaling: B per GPU)
    host_batch = synthetic_events(args.workload, B, seed=1234 + rank)
    x0_host = synthetic_noise(host_batch, seed=rank)
    counts = host_batch["q_mask"].sum(1).numpy()
    T = int(counts.sum())
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host_batch.items()}
    x0_pinned = x0_host.pin_memory()

    lib = _lib.load()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident leg
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host_batch.items()}
    ev = model._bind(dbatch)
    h = model._handle
    x0p = ev.pack(x0_host.to(dev))
    out = torch.empty(T, dtype=torch.float32, device=dev)
    tgrid = torch.linspace(0, 1, args.n_steps).float().contiguous()
    nfe = C.c_int32(0)

    def sample_resident():
        rc = lib.srhep_sample(h, x0p.data_ptr(), tgrid.data_ptr(), args.n_steps, _lib.METHODS["euler"], 0, out.data_ptr(), C.byref(nfe), stream)
        _lib.check(lib, h, rc, "srhep_sample")

    for _ in range(max(args.warmup, 3)):
        sample_resident()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = model.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sample_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count - l0
    clk = clocks.stop() if rank == 0 else None
    if not torch.isfinite(out).all():
        raise SystemExit("non-finite samples")
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ------------------------------------------------------------------ end-to-end leg
    h2d = sum(v.numel() * v.element_size() for v in host_batch.values() if torch.is_tensor(v))
    d2h = x0_host.numel() * 4
    result_host = torch.empty(x0_host.shape, dtype=torch.float32).pin_memory()
    from superresolutionhep_b200 import sharding

    def step_e2e():
        b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
        x1 = model.generate_samples(b, n_steps=args.n_steps, method="euler")          # noise drawn on device, as the reference does
        if world > 1:                                                                  # the one collective: packed final outputs over NVLink
            sharding.gather_packed(x1[..., 0][b["q_mask"]], counts, dst=0)
        result_host.copy_(x1, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(max(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    t2 = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, wall)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t2.item())

    # ------------------------------------------------------------------ roofline leg (rank 0)
    line = None
    if rank == 0:
        peaks = load_peaks()
        ev = model._bind(dbatch)
        v = torch.empty(T, dtype=torch.float32, device=dev)
        ms_cat = (C.c_float * len(_lib.CATEGORIES))()
        n_cat = (C.c_int32 * len(_lib.CATEGORIES))()
        for _ in range(2):
            _lib.check(lib, model._handle, lib.srhep_profile(model._handle, x0p.data_ptr(), 0.5, v.data_ptr(), ms_cat, n_cat, stream), "srhep_profile")
        per_cat = {c: {"ms": float(ms_cat[i]), "launches": int(n_cat[i])} for i, c in enumerate(_lib.CATEGORIES)}
        n = counts.astype(np.float64)
        H, L = model.dims.h_dim, model.dims.layers
        unit = 2.0 * H * H * n.sum()                                     # one 256x256 Linear over every real cell
        fused = per_cat.get("chain", {"launches": 0})["launches"] > 0   # layer chain kernel in use (kernels_chain.cuh)
        algo = {                                                         # algorithmic FLOPs of one evaluation per category (real cells only)
            "attn": 4.0 * H * (n * n).sum() * L,
            "qkv": 3 * unit * (1 if fused else L), "out": 0.0 if fused else unit * L, "mlp1": 0.0 if fused else unit * L,
            "mlp2": 0.0 if fused else unit * L, "chain": unit * (3 * L + 3 * (L - 1)) if fused else 0.0,
        }
        total_ms = sum(c["ms"] for c in per_cat.values())
        dom = max(algo, key=lambda c: per_cat[c]["ms"])
        dom_ms = per_cat[dom]["ms"]
        dom_launch = max(per_cat[dom]["launches"], 1)
        achieved = algo[dom] / dom_launch / (dom_ms / dom_launch * 1e-3) / 1e12
        peak = peaks["tf_sustained"]
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": load_traffic(args.workload, B, args.precision).get(dom), "peak_source": f"{peaks['source']} (sustained bf16, kernel timed inside a long step)",
                "kernel_ms_per_launch": dom_ms / dom_launch, "algorithmic_flops_per_launch": algo[dom] / dom_launch, "share_of_evaluation": dom_ms / total_ms if total_ms else None,
                "per_category_ms": {c: round(per_cat[c]["ms"], 4) for c in per_cat},
                "evaluation_tflops": flops_per_eval(counts) / (total_ms * 1e-3) / 1e12 if total_ms else None}
        cpu = None
        if not args.no_cpu_baseline and world == 1:          # reported at N = 1 only (under torchrun the host threads are pinned to 1 per rank)
            cores = os.cpu_count() or 1
            val, secs = oracle_events_per_s(args.workload, args.ref_events, args.n_steps)
            cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.ref_events} {args.workload} events, euler n_steps={args.n_steps}, fp32 torch CPU, {secs:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": workload_config(args, B, args.precision),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "cells_per_gpu": T, "evaluations_per_step": int(nfe.value),
            "tflops_algorithmic": flops_per_eval(counts) * nfe.value * args.steps * world / (ms_max * 1e-3) / 1e12,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _emit(line)


if __name__ == "__main__":
    main()
