#!/usr/bin/env python
"""Where the time of one end-to-end ``generate_samples`` call goes (host batch in pinned memory, 4096 single_e events):
H2D, packing, binding (host planning + graph), sampling, unpacking, D2H.  Every phase is bracketed by a device
synchronise, so the phases add up to a little more than the un-instrumented call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolutionhep_b200 import FlowModel, _lib
from superresolutionhep_b200.flow_model import PackedEvents
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
m = FlowModel(flow_config("single_e"), precision="bf16"); m.load_state_dict(synthetic_state_dict(m.dims, seed=7)); m.eval().cuda()
host = synthetic_events("single_e", B, seed=1234)
pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
out_host = torch.empty(host["e_proxy"].shape, dtype=torch.float32).pin_memory()

def sync(): torch.cuda.synchronize()
def timed(f):
    sync(); t = time.perf_counter(); r = f(); sync(); return r, (time.perf_counter() - t) * 1e3

for it in range(4):
    b, t_h2d = timed(lambda: {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()})
    _, t_pack = timed(lambda: PackedEvents(b, dev))
    _, t_bind = timed(lambda: m._bind(b))                      # includes a second PackedEvents
    x0, t_noise = timed(lambda: torch.randn_like(b["e_proxy"]))
    x1, t_samp = timed(lambda: m.generate_samples(b, n_steps=25, method="euler", x0=x0))     # bound already: pack x0 + sample + unpack
    _, t_d2h = timed(lambda: out_host.copy_(x1, non_blocking=True))
    _, t_all = timed(lambda: out_host.copy_(m.generate_samples({k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()},
                                                                 n_steps=25, method="euler"), non_blocking=True))
    print(f"iter {it}: h2d {t_h2d:.2f}  PackedEvents {t_pack:.2f}  bind(total) {t_bind:.2f}  noise {t_noise:.2f}  sample(bound) {t_samp:.2f}  d2h {t_d2h:.2f}  | whole call {t_all:.2f} ms")
