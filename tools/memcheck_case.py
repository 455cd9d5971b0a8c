#!/usr/bin/env python
"""Small ragged case for compute-sanitizer: 300 single_e events (row count not a multiple of 128, more than 256 events so the
128x128 fp32 GEMM tile is used), bf16 sampling (shared-time path) and one forward (per-event-time path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolutionhep_b200 import FlowModel
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict
for kind, B in (("single_e", 300), ("multipart", 40)):
    m = FlowModel(flow_config(kind), precision="bf16"); m.load_state_dict(synthetic_state_dict(m.dims, seed=7)); m.eval().cuda()
    b = synthetic_events(kind, B, seed=5); x = synthetic_noise(b, seed=1)
    db = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    xs = m.generate_samples(db, n_steps=3, method="midpoint", ret_seq=True, x0=x.cuda())
    v = m(db, x.cuda(), torch.rand(B).cuda())
    torch.cuda.synchronize()
    print(kind, "cells", int(b["q_mask"].sum()), "finite", bool(torch.isfinite(xs[:, b["q_mask"]]).all()), bool(torch.isfinite(v[b["q_mask"]]).all()))
