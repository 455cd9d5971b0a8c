#!/usr/bin/env python
"""Benchmark of the SR sampling hot path (BASELINE.json metric: events/s of SR sampling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload single_e|multipart] [--events B] [--n-steps S] [--precision fp16|bf16|fp32]
                    [--no-extra] [--sweep]

One "step" = one ``generate_samples`` pass (Euler, ``n_steps`` grid points = ``n_steps-1``
network evaluations) over one batch of ``B`` synthetic events per GPU.  Default workload =
BASELINE.json configs[1]: single-electron shapes, B = 4096 per GPU, n_steps = 25, 16-bit tcgen05
operands (fp16 by default: the 16-bit format that meets the north-star tolerance per evaluation).

* ``value``  : whole-job events/s with packed inputs resident in HBM, timed with CUDA events
               around K calls of the C-ABI ``srhep_sample`` (max over ranks).
* ``e2e``    : the same metric through the public API (``FlowModel.generate_samples``) with
               the batch in pinned HOST memory: H2D of the batch, packing, binding, sampling,
               D2H of the final cells all inside the timed region (plus, for N > 1, the final
               NCCL gather of the outputs to rank 0).
* ``roofline``: dominant kernel category of one evaluation, timed live with CUDA events on
               the launching stream (srhep_profile), against MEASURED_PEAKS.json.
* ``cpu_baseline``: the reference's own modules (oracle/_ref, see oracle/make_ref.py; the oracle port
               when that copy is absent) on a bounded sample of the same workload, all host threads.
* ``--impl reference``: the reference arm = the reference's ``FlowModel.generate_samples`` timed as the whole run.

Further legs on the same JSON line (``--no-extra`` skips them; none of them changes ``value``):
* ``multipart_strong`` (BASELINE configs[2]): ONE shared list of 16 384 multi-particle events cut into cost-balanced
               entry ranges (sharding.plan_entry_ranges), one range per rank, outputs gathered on rank 0; reports
               events/s, the per-rank times (imbalance = max / mean) and, at N > 1, the efficiency against rank 0
               sampling the same list alone.
* ``multipart_sweep`` (configs[3], N > 1 or ``--sweep``): the same sharded list at n_steps 10 / 25 / 50 / 100.
* ``pflow_sharded`` (configs[4]): SAPF forward with the real pf_hr weights over one shared list of SR-output-like events
               sharded by sharding.PFLOW_COST.
* ``dopri5`` (N = 1): the reference's default solver (adaptive, atol = rtol = 1e-4) on the device, batch 20 and the
               full batch, with the number of function evaluations; in the benchmarked operand mode and in 'highest' (fp32-grade
               on the tensor cores), which together with dopri5 is what ``inference.py`` runs when no flag is given.
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
    p = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if not os.path.isfile(p):
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


def _cpu_sampler(kind: str):
    """(sample(batch, x0, n_steps) -> x1, kind) on the CPU: the reference's OWN ``FlowModel.generate_samples`` when its sources are
    importable (the mount in the build container, the oracle/_ref copy on the GPU box), driven by the restated fixed-grid
    ``odeint`` (torchdiffeq itself is not installable); the oracle port otherwise."""
    from oracle import odeint as _ode, ref_import, sr_oracle
    from superresolutionhep_b200.config import SrDims
    cfg = flow_config(kind)
    sd = synthetic_state_dict(SrDims.from_config(cfg), seed=WEIGHT_SEED)
    if ref_import.available():
        m = ref_import.build_reference_flow_model(cfg, sd)
        sys.modules["torchdiffeq"].odeint = _ode.odeint                      # the one missing third-party piece (stub module of ref_import)

        def sample(batch, x0, n_steps):
            torch.manual_seed(0)
            return m.generate_samples(batch, n_steps=n_steps, method="euler")      # reference API, reference code path (draws its own noise)
        return sample, "reference", f"reference FlowModel.generate_samples ({ref_import.source()})"
    dims = sr_oracle.derive_dims(cfg)
    return (lambda batch, x0, n_steps: sr_oracle.generate_samples(sd, dims, batch, x0, n_steps=n_steps, method="euler")), "port", "oracle port"


def oracle_events_per_s(kind: str, n_events: int, n_steps: int, repeats: int = 1):
    """The CPU implementation on a bounded sample: ``n_events`` events, one padded batch, Euler."""
    sample, how, what = _cpu_sampler(kind)
    batch = synthetic_events(kind, n_events, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    torch.set_num_threads(os.cpu_count() or 1)
    best = None
    with torch.no_grad():
        sample(batch, x0, 2)                                                 # warm-up evaluation
        for _ in range(repeats):
            t0 = time.perf_counter()
            sample(batch, x0, n_steps)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_events / best, best, how, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ev = args.ref_events
    times = []
    sample, how, what = _cpu_sampler(args.workload)
    batch = synthetic_events(args.workload, n_ev, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            sample(batch, x0, args.n_steps)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    val = n_ev * len(times) / total
    sample_txt = (f"{n_ev} synthetic {args.workload} events per step (one padded batch), euler n_steps={args.n_steps} "
                  f"({args.n_steps - 1} evaluations), fp32, {what}, torch CPU {torch.__version__}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, args.events, args.precision), "reference_sample": sample_txt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": how, "sample": sample_txt},
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
    ap.add_argument("--precision", default=os.environ.get("SRHEP_BENCH_PRECISION", "fp16"), choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--ref-events", type=int, default=64, help="events per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--pass-tokens", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the multipart_strong / sweep / pflow / dopri5 legs")
    ap.add_argument("--sweep", action="store_true", help="run the n_steps sweep at N = 1 too")
    ap.add_argument("--strong-events", type=int, default=16384, help="events of the shared multipart list (strong-scaling leg)")
    ap.add_argument("--pflow-events", type=int, default=16384, help="events of the shared pflow list")
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
            val, secs, how, what = oracle_events_per_s(args.workload, args.ref_events, args.n_steps)
            cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": how,
                   "sample": f"{args.ref_events} {args.workload} events, euler n_steps={args.n_steps}, fp32 torch CPU, {what}, {secs:.1f} s"}
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
    if not args.no_extra and args.precision != "fp32":
        model.release()
        del model, dbatch, ev, x0p, out
        torch.cuda.empty_cache()
        extra = extra_legs(args, world, rank, dev)
        if rank == 0:
            line.update(extra)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        _emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] and the reference's default solver, on the same JSON line
# ---------------------------------------------------------------------------------------------------------------------
def _max_over_ranks(x: float, dev, world: int) -> float:
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _all_ranks(x: float, dev, world: int):
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world == 1:
        return [float(x)]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def extra_legs(args, world: int, rank: int, dev) -> dict:
    """Every leg is fenced: a failure is reported under its key and never costs the headline line.  (The legs run the same
    code on every rank, so a deterministic failure is raised by all ranks at the same point and no collective is left
    half-entered.)"""
    import traceback
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def fenced(fn, *a):
        try:
            return fn(*a)
        except Exception as e:                                            # noqa: BLE001
            traceback.print_exc()
            torch.cuda.empty_cache()
            return {"error": f"{type(e).__name__}: {e}"[:400]}

    res = {}
    mp = fenced(multipart_legs, args, world, rank, dev, barrier)
    if "error" in mp:
        res["multipart_strong"] = mp
    else:
        res.update(mp)
    res["pflow_sharded"] = fenced(pflow_leg, args, world, rank, dev, barrier)
    if world == 1:
        res["dopri5"] = fenced(dopri5_leg, args, dev)
    return res


def multipart_legs(args, world: int, rank: int, dev, barrier) -> dict:
    import torch.distributed as dist
    from superresolutionhep_b200 import FlowModel, sharding
    res = {}
    # ---------------------------------------------------------------- multipart, strong scaling over one shared list (configs[2])
    N = args.strong_events
    cfg = flow_config("multipart")
    model = FlowModel(cfg, precision=args.precision)
    model.load_state_dict(synthetic_state_dict(model.dims, seed=WEIGHT_SEED))
    model.eval().cuda(dev)
    full = synthetic_events("multipart", N, seed=4321)                  # every rank builds the SAME list
    counts = full["q_mask"].sum(1).numpy()
    x0_full = synthetic_noise(full, seed=11)
    ranges = sharding.plan_entry_ranges(counts, world)
    a, b = ranges[rank]
    cost = sharding.event_cost(counts)

    def shard(a, b):
        sub = sharding.shard_batch(full, a, b)
        sub = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sub.items()}
        return sub, x0_full[a:b, : sub["q_mask"].shape[1]].to(dev)

    sub, x0_sub = shard(a, b)

    def run_sharded(n_steps, gather=True):
        """-> (whole-job seconds incl. the final gather, this rank's sampling seconds)"""
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        e0.record()
        xs = model.generate_samples(sub, n_steps=n_steps, method="euler", x0=x0_sub) if b > a else None
        e1.record()
        if gather:
            packed = xs[..., 0][sub["q_mask"]] if xs is not None else torch.zeros(0, device=dev)
            sharding.gather_packed(packed, counts[a:b], dst=0)
        e2.record()
        torch.cuda.synchronize(dev)
        return _max_over_ranks(e0.elapsed_time(e2) * 1e-3, dev, world), e0.elapsed_time(e1) * 1e-3

    run_sharded(3)                                                       # warm-up: workspace allocation, graph capture, NCCL channels
    tot, mine = run_sharded(args.n_steps)
    tot2, mine2 = run_sharded(args.n_steps)
    if tot2 < tot:
        tot, mine = tot2, mine2
    per_rank = _all_ranks(mine, dev, world)
    strong = {"workload": f"multipart SR sampling, ONE shared list of {N} synthetic events cut into {world} cost-balanced entry ranges, euler n_steps={args.n_steps}",
              "events": N, "cells": int(counts.sum()), "events_s": N / tot, "seconds": tot, "per_rank_seconds": [round(t, 4) for t in per_rank],
              "imbalance": max(per_rank) / (sum(per_rank) / len(per_rank)), "planned_cost_share": [float(cost[x:y].sum() / cost.sum()) for x, y in ranges],
              "events_per_rank": [y - x for x, y in ranges],
              "tflops_algorithmic": flops_per_eval(counts) * (args.n_steps - 1) / tot / 1e12, "efficiency": None, "single_gpu_events_s": None}
    if world > 1:                                                        # the N = 1 run of the same list, on rank 0 alone
        t1 = None
        if rank == 0:
            try:
                whole, x0_whole = shard(0, N)
                model.generate_samples(whole, n_steps=3, method="euler", x0=x0_whole)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                e0.record()
                model.generate_samples(whole, n_steps=args.n_steps, method="euler", x0=x0_whole)
                e1.record()
                torch.cuda.synchronize(dev)
                t1 = e0.elapsed_time(e1) * 1e-3
                del whole, x0_whole
            except Exception as e:                                        # noqa: BLE001  (the other ranks wait at the barrier below)
                strong["single_gpu_error"] = f"{type(e).__name__}: {e}"[:300]
        barrier()
        if rank == 0 and t1:
            strong["single_gpu_events_s"] = N / t1
            strong["efficiency"] = t1 / (world * tot)
    else:
        strong["single_gpu_events_s"] = strong["events_s"]
        strong["efficiency"] = 1.0
    res["multipart_strong"] = strong
    # ---------------------------------------------------------------- n_steps sweep on the same sharded list (configs[3])
    if world > 1 or args.sweep:
        sweep = {}
        for ns in (10, 25, 50, 100):
            tot_s, _ = run_sharded(ns)
            sweep[str(ns)] = {"events_s": N / tot_s, "seconds": tot_s, "evaluations": ns - 1,
                              "tflops_algorithmic": flops_per_eval(counts) * (ns - 1) / tot_s / 1e12}
        res["multipart_sweep"] = {"workload": f"the multipart_strong list ({N} events over {world} GPUs), euler, one CUDA graph per pass replayed per step", "n_steps": sweep}
    del sub, x0_sub, full, x0_full
    model.release()
    torch.cuda.empty_cache()
    return res


def pflow_leg(args, world, rank, dev, barrier) -> dict:
    import torch.distributed as dist
    from superresolutionhep_b200 import sharding
    from superresolutionhep_b200.pflow import PflowLightning
    from superresolutionhep_b200.synthetic import synthetic_pflow_events
    g = torch.load(os.path.join(ROOT, "tests", "golden", "pflow_pf_hr.pt"))
    lm = PflowLightning({"pf_model": g["pf_model"], "var_transform": g["var_transform"]}, {}, inference=True)
    lm.load_state_dict({"net." + k: v for k, v in g["state_dict"].items()}, strict=True)
    lm.eval().cuda(dev)
    N = args.pflow_events
    full = synthetic_pflow_events(N, seed=5)
    counts = full["cell_mask"].sum(1).numpy()
    ranges = sharding.plan_entry_ranges(counts, world, sharding.PFLOW_COST)
    a, b = ranges[rank]
    sub = sharding.shard_batch(full, a, b, mask_key="cell_mask")
    sub = {k: v.to(dev) for k, v in sub.items()}

    def once():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        e0.record()
        logits, kin, inc = lm.net(sub)
        e1.record()
        if world > 1:                                                    # the event-level outputs of every range to rank 0 (entry order)
            pay = torch.cat([logits, kin.reshape(kin.shape[0], -1)], 1).t().contiguous()       # (21, B_r)
            sharding.gather_packed(pay, np.ones(pay.shape[1], dtype=np.int64), dst=0)
        e2.record()
        torch.cuda.synchronize(dev)
        return _max_over_ranks(e0.elapsed_time(e2) * 1e-3, dev, world), e0.elapsed_time(e1) * 1e-3

    once(); once()
    best, mine = min(once() for _ in range(3))
    per_rank = _all_ranks(mine, dev, world)
    n = counts.astype(np.float64)
    return {"workload": f"SAPF forward (real pf_hr weights), ONE shared list of {N} synthetic SR-output events cut into {world} ranges balanced on {sharding.PFLOW_COST}",
            "events": N, "cells": int(counts.sum()), "events_s": N / best, "seconds": best, "per_rank_seconds": [round(t, 4) for t in per_rank],
            "imbalance": max(per_rank) / (sum(per_rank) / len(per_rank)), "events_per_rank": [y - x for x, y in ranges],
            "tflops_algorithmic": float((214e3 * n + 768.0 * n * n).sum()) / best / 1e12,
            "note": "packing of the padded batch and unpacking of inc_weights (PyTorch indexing) are inside the timed region"}


def dopri5_leg(args, dev) -> dict:
    """The reference's real default: ``inference.py -p highest`` (fp32) + ``method="dopri5"`` (models/flow_model.py:303).  Both the
    benchmarked 16-bit operand mode and the fp32-grade tensor-core mode ('highest') are timed."""
    from superresolutionhep_b200 import FlowModel
    cfg = flow_config(args.workload)
    out = {"workload": f"{args.workload} SR sampling with the reference's default solver: dopri5, atol = rtol = 1e-4, n_steps={args.n_steps} output grid points; "
                       "the adaptive loop runs on the device as one conditional CUDA graph", "runs": []}
    for prec in (args.precision, "fp32"):
        model = FlowModel(cfg, precision=prec)
        model.load_state_dict(synthetic_state_dict(model.dims, seed=WEIGHT_SEED))
        model.eval().cuda(dev)
        for B in (20, args.events):                                      # the shipped batch size (configs/single_e/inference_batch.yml:3) and the full batch
            batch = synthetic_events(args.workload, B, seed=1234)
            x0 = synthetic_noise(batch, seed=0).to(dev)
            db = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
            model.generate_samples(db, n_steps=args.n_steps, method="dopri5", x0=x0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 if prec != "fp32" or B <= 64 else 1
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(reps):
                model.generate_samples(db, n_steps=args.n_steps, method="dopri5", x0=x0)
            e1.record()
            torch.cuda.synchronize(dev)
            sec = e0.elapsed_time(e1) * 1e-3 / reps
            out["runs"].append({"precision": prec + (" ('highest': fp32-grade on the tensor cores, hi/lo fp16 operand planes)" if prec == "fp32" else ""),
                                "events": B, "events_s": B / sec, "seconds": sec, **model.last_stats})
        model.release()
        del model
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
