"""Per-stage error of the bf16 path against the fp32 oracle (diagnostic, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sr_oracle
from superresolutionhep_b200 import FlowModel
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict

kind = sys.argv[1] if len(sys.argv) > 1 else "single_e"
counts = np.array([int(c) for c in sys.argv[2].split(",")]) if len(sys.argv) > 2 else np.array([4, 128, 132, 36, 260, 500])
cfg = flow_config(kind)
m = FlowModel(cfg, precision=os.environ.get("DIAG_PREC", "bf16"))
sd = synthetic_state_dict(m.dims, seed=21)
m.load_state_dict(sd); m.cuda()
dims = sr_oracle.derive_dims(cfg)
batch = synthetic_events(kind, len(counts), seed=3, counts=counts)
x = synthetic_noise(batch, seed=4)
t = torch.linspace(0.0, 1.0, len(counts))
taps = {}
with torch.no_grad():
    ref = sr_oracle.flow_forward(sd, dims, batch, x, t, taps=taps)
names = ["time_emb", "context", "feat_0"] + [f"layer_{i}" for i in range(m.dims.layers)] + ["transformer_out"]
dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
got, ev = m.debug_taps(dev, x.cuda(), t.cuda(), names)
mask = batch["q_mask"]
for n in names:
    r = taps[n] if n in ("time_emb", "context") else taps[n][mask]
    g = got[n].cpu()
    err = (g - r).abs().max().item(); scale = r.abs().max().item()
    rel2 = ((g - r).norm() / r.norm()).item()
    print(f"{n:16s} max|err| {err:.4e}  max|ref| {scale:.3e}  relL2 {rel2:.3e}  finite {bool(torch.isfinite(g).all())}")
v = got["v_t"].cpu()[mask]; r = ref[mask]
print(f"v_t              max|err| {(v-r).abs().max().item():.4e}  max|ref| {r.abs().max().item():.3e}  relL2 {((v-r).norm()/r.norm()).item():.3e}")
