#!/usr/bin/env python
"""events/s of the particle-flow forward (BASELINE.json configs[4]) on synthetic SR-output cells, one GPU:
real pf_hr weights (tests/golden/pflow_pf_hr.pt), cells packed on the device, CUDA-event timing, CPU oracle beside it."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superresolutionhep_b200.pflow import PflowLightning                     # noqa: E402
from superresolutionhep_b200.synthetic import synthetic_pflow_events         # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.load(os.path.join(ROOT, "tests", "golden", "pflow_pf_hr.pt"))
lm = PflowLightning({"pf_model": g["pf_model"], "var_transform": g["var_transform"]}, {}, inference=True)
lm.load_state_dict({"net." + k: v for k, v in g["state_dict"].items()}, strict=True)
lm.eval().cuda()
batch = {k: v.cuda() for k, v in synthetic_pflow_events(B, seed=5).items()}
cells = int(batch["cell_mask"].sum())
for _ in range(3):
    out = lm.net(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
l0 = lm.net.launch_count
e0.record()
for _ in range(K):
    out = lm.net(batch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
# CPU oracle on a bounded sample
from oracle import pflow_oracle                                               # noqa: E402
cb = synthetic_pflow_events(64, seed=6)
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], cb)
    t0 = time.perf_counter(); pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], cb); dt = time.perf_counter() - t0
n = batch["cell_mask"].sum(1).double()
flops = float((214e3 * n + 768.0 * n * n).sum())
print(json.dumps({"metric": "events/sec pflow forward (SAPF, pf_hr)", "value": B / (ms * 1e-3), "unit": "events/s", "events": B, "cells": cells, "ms_per_forward": ms,
                  "gpu_launches_per_forward": (lm.net.launch_count - l0) // K, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                  "note": "includes packing the padded batch (mask indexing) and unpacking inc_weights in PyTorch",
                  "cpu_baseline": {"value": 64 / dt, "unit": "events/s", "cores": os.cpu_count(), "kind": "port", "sample": "64 events"}}))
