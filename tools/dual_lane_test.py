#!/usr/bin/env python
"""Experiment: K independent lanes (handles) on K streams, each sampling B / K events, with the persistent chain / attention
grids at 1 or 2 CTAs per SM.  Question: does an attention CTA (SFU-bound) next to a chain CTA (latency / L2-bound) on the
same SM beat two CTAs of the same kernel?   python tools/dual_lane_test.py [events] [n_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolutionhep_b200 import FlowModel
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict
from superresolutionhep_b200 import sharding

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kind = sys.argv[3] if len(sys.argv) > 3 else "single_e"
cfg = flow_config(kind)
full = synthetic_events(kind, B, seed=1234)
x0 = synthetic_noise(full, seed=0)
counts = full["q_mask"].sum(1).numpy()
dev = torch.device("cuda", 0)


def run(lanes, ctas, delay_ms, reps=3):
    os.environ["SRHEP_CTAS_PER_SM"] = str(ctas)
    ranges = sharding.plan_entry_ranges(counts, lanes)
    models, subs, xs, streams = [], [], [], []
    for a, b in ranges:
        m = FlowModel(cfg, precision="fp16"); m.load_state_dict(synthetic_state_dict(m.dims, seed=7)); m.eval().cuda(dev)
        sub = sharding.shard_batch(full, a, b)
        sub = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sub.items()}
        models.append(m); subs.append(sub); xs.append(x0[a:b, : sub["q_mask"].shape[1]].to(dev)); streams.append(torch.cuda.Stream(dev))
    outs = [None] * lanes

    def once():
        for i in range(lanes):
            with torch.cuda.stream(streams[i]):
                if i and delay_ms:
                    torch.cuda._sleep(int(i * delay_ms * 1.7e6))
                outs[i] = models[i].generate_samples(subs[i], n_steps=NS, method="euler", x0=xs[i])
    once(); torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in streams: s.wait_event(e0)
        once()
        for s in streams: e1.wait(s) if False else torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    chk = float(sum(o.double().abs().sum() for o in outs))
    for m in models: m.release()
    print(f"lanes={lanes} ctas_per_sm={ctas} delay={delay_ms} ms: {best:.1f} ms  {B / best * 1e3:.0f} events/s  checksum {chk:.6e}", flush=True)


run(1, 2, 0)
run(2, 1, 0)
run(2, 1, 0.5)
run(2, 1, 1.0)
run(2, 1, 1.5)
run(2, 2, 0)
run(2, 2, 1.0)
run(3, 1, 0.7)
run(1, 1, 0)
