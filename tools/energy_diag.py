#!/usr/bin/env python
"""Energy per evaluation by kernel group (diagnostic; results of the SRHEP_ONLY runs are wrong on purpose).

The step sits at the 1 000 W power cap, so what a kernel costs is its ENERGY, not its time at some clock.  This tool replays,
for a few seconds each, (0) whole evaluations, (1) the six attention launches alone, (2) the six layer-chain launches alone,
(3) everything else, samples nvidia-smi power / SM clock meanwhile and prints ms, W, MHz and J per evaluation for each group.
    python tools/energy_diag.py [events] [workload]
"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from superresolutionhep_b200 import FlowModel
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else "single_e"
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
m = FlowModel(flow_config(kind), precision=prec); m.load_state_dict(synthetic_state_dict(m.dims, seed=7)); m.eval().cuda()
m.use_graph = False                                    # the switch is read per API call; captured graphs would replay the old schedule
b = synthetic_events(kind, B, seed=1234); x = synthetic_noise(b, seed=0)
db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
x0 = x.cuda()


class Sampler:
    def __init__(self):
        self.rows, self.proc = [], None
    def start(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, text=True)
        threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True).start()
    def stop(self):
        time.sleep(0.15); self.proc.terminate()
        pw, ck = [], []
        for l in self.rows[3:]:                      # the first samples still see the previous phase
            try:
                a, c = l.split(","); pw.append(float(a)); ck.append(float(c))
            except ValueError:
                pass
        return (float(np.median(pw)) if pw else float("nan")), (float(np.median(ck)) if ck else float("nan"))


def run(only, n_steps, secs=4.0):
    os.environ.pop("SRHEP_ONLY", None)
    m.generate_samples(db, n_steps=3, method="euler", x0=x0)          # realistic buffer contents
    if only:
        os.environ["SRHEP_ONLY"] = str(only)
    m.generate_samples(db, n_steps=3, method="euler", x0=x0)
    torch.cuda.synchronize()
    s = Sampler(); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); evals = 0
    e0.record()
    while time.time() - t0 < secs:
        m.generate_samples(db, n_steps=n_steps, method="euler", x0=x0)
        evals += n_steps - 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / evals
    pw, ck = s.stop()
    os.environ.pop("SRHEP_ONLY", None)
    return ms, pw, ck


names = {0: "whole evaluation", 1: "attention launches only", 2: "layer-chain launches only", 3: "everything else"}
idle = None
time.sleep(1.0)
s = Sampler(); s.start(); time.sleep(1.5); idle, ick = s.stop()
print(f"idle: {idle:.0f} W at {ick:.0f} MHz")
for only in (0, 1, 2, 3, 0):
    ms, pw, ck = run(only, 25 if only in (0, 2) else 49)
    print(f"{names[only]:28s} {ms:7.3f} ms/eval  {pw:6.0f} W  {ck:5.0f} MHz  {ms * 1e-3 * pw:6.2f} J/eval  ({ms * 1e-3 * (pw - idle):6.2f} J above idle)", flush=True)
