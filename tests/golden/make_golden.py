"""Mint golden vectors from the UNMODIFIED reference modules (run in the build container).

    python tests/golden/make_golden.py

Needs /root/reference (read-only mount).  The velocity network is the reference's own
``FlowModel`` (imported with the two stub modules of oracle/ref_import.py); the ODE driver is
oracle/odeint.py (torchdiffeq is not installable, see its header).  Weights come from
``synthetic_state_dict`` (seeded) because the SR checkpoints are missing from the mount.
Outputs (committed): tests/golden/sr_taps_single_e.pt, sr_taps_multipart.pt,
sr_config1_single_e.pt, sr_dopri5_single_e.pt, sr_traj_multipart.pt.

    python tests/golden/make_golden.py [case ...]      # taps | dopri5 | config1 | multipart (default: all)
"""
import os
import sys
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import                       # noqa: E402
from oracle.odeint import odeint                    # noqa: E402
from superresolutionhep_b200.config import SrDims   # noqa: E402
from superresolutionhep_b200.synthetic import (     # noqa: E402
    synthetic_events, synthetic_noise, synthetic_state_dict)

OUT = os.path.dirname(os.path.abspath(__file__))
WEIGHT_SEED = 7


def load_cfg(name):
    with open(f"/root/reference/configs/{name}/model_and_var.yml") as fp:
        return yaml.safe_load(fp)["flow_model"]


def ref_model(name):
    cfg = load_cfg(name)
    sd = synthetic_state_dict(SrDims.from_config(cfg), seed=WEIGHT_SEED)
    return ref_import.build_reference_flow_model(cfg, sd)


def taps_case(name, counts, pad_to, fname):
    """Per-stage activations via forward hooks on the reference modules."""
    m = ref_model(name)
    batch = synthetic_events(name, len(counts), seed=11, counts=np.array(counts), pad_to=pad_to)
    x = synthetic_noise(batch, seed=5)
    t = torch.linspace(0.05, 0.95, len(counts))
    taps = {}
    hooks = [
        m.time_step_embedder.register_forward_hook(lambda mod, i, o: taps.__setitem__("time_emb", o.clone())),
        m.feat_0_mlp.register_forward_hook(lambda mod, i, o: taps.__setitem__("feat_0", o.clone())),
        m.transformer.layers[0].register_forward_hook(lambda mod, i, o: taps.__setitem__("layer_0", o.clone())),
        m.transformer.layers[-1].register_forward_hook(lambda mod, i, o: taps.__setitem__(f"layer_{len(m.transformer.layers)-1}", o.clone())),
        m.transformer.register_forward_hook(lambda mod, i, o: taps.__setitem__("transformer_out", o.clone())),
        m.feat_0_mlp.register_forward_hook(lambda mod, i, o: taps.__setitem__("context", i[1].clone()) if len(i) > 1 else None),
    ]
    # feat_0_mlp is called as feat_0_mlp(feat_0, context=context): context arrives as a kwarg
    hooks.append(m.feat_0_mlp.register_forward_pre_hook(
        lambda mod, args, kwargs: taps.__setitem__("context", kwargs["context"].clone()), with_kwargs=True))
    with torch.no_grad():
        v = m(batch, x, t)
    for h in hooks:
        h.remove()
    taps["v_t"] = v
    torch.save({"config": name, "weight_seed": WEIGHT_SEED, "counts": list(counts), "pad_to": pad_to,
                "event_seed": 11, "noise_seed": 5, "t": t, "taps": taps}, os.path.join(OUT, fname))
    print(fname, {k: tuple(v.shape) for k, v in taps.items()})


def config1_case(fname, n_events=64, n_steps=25):
    """BASELINE.json configs[0]: 64 single_e events, fixed noise seed, Euler + midpoint."""
    m = ref_model("single_e")
    batch = synthetic_events("single_e", n_events, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    B = x0.shape[0]
    out = {"config": "single_e", "weight_seed": WEIGHT_SEED, "event_seed": 1234, "noise_seed": 0,
           "n_events": n_events, "n_steps": n_steps}
    for method in ("euler", "midpoint"):
        rec = []

        def f(t, x):
            v = m(batch, x, t * torch.ones(B))
            rec.append(v.clone())
            return v
        t0 = time.time()
        with torch.no_grad():
            xs = odeint(f, x0, torch.linspace(0, 1, n_steps), method=method)
        print(method, "evals", len(rec), f"{time.time()-t0:.1f}s")
        keep = [0, len(rec) // 2, len(rec) - 1]
        out[method] = {"x_final": xs[-1].clone(), "x_mid": xs[n_steps // 2].clone(),
                       "v_evals": {k: rec[k] for k in keep}, "nfe": len(rec)}
    torch.save(out, os.path.join(OUT, fname))


def dopri5_case(fname, n_events=6, n_steps=5):
    m = ref_model("single_e")
    batch = synthetic_events("single_e", n_events, seed=77)
    x0 = synthetic_noise(batch, seed=3)
    B = x0.shape[0]
    stats = {}
    with torch.no_grad():
        xs = odeint(lambda t, x: m(batch, x, t * torch.ones(B)), x0, torch.linspace(0, 1, n_steps),
                    method="dopri5", atol=1e-4, rtol=1e-4, stats=stats)
    print("dopri5", stats)
    torch.save({"config": "single_e", "weight_seed": WEIGHT_SEED, "event_seed": 77, "noise_seed": 3,
                "n_events": n_events, "n_steps": n_steps, "x_seq": xs, "stats": stats},
               os.path.join(OUT, fname))


MULTIPART_COUNTS = [3280, 2048, 1600, 1312, 960, 800, 640, 640, 480, 400, 320, 256, 128, 48, 16, 16]


def multipart_case(fname, counts=MULTIPART_COUNTS, n_steps=25, group=4):
    """BASELINE.json configs[2] shapes: 16 multipart events from 16 cells up to the maximum of 3280, Euler + midpoint
    trajectories with n_steps = 25.  Events are independent under a fixed grid, so the reference model is run on groups of
    `group` events of similar length (padding all 16 to 3280 cells would cost 4x the time for the same real rows); the
    noise of event i is row i of ONE seeded (B, Nmax, 1) draw, so a single-batch run of the product sees the same x0.
    Stored packed (real cells only, entry order)."""
    m = ref_model("multipart")
    counts = np.asarray(counts)
    batch = synthetic_events("multipart", len(counts), seed=4242, counts=counts)
    x0 = synthetic_noise(batch, seed=17)
    mask = batch["q_mask"]
    out = {"config": "multipart", "weight_seed": WEIGHT_SEED, "event_seed": 4242, "noise_seed": 17, "counts": [int(c) for c in counts],
           "n_steps": n_steps}
    for method in ("euler", "midpoint"):
        finals, mids, v0s, nfe = [], [], [], 0
        t0 = time.time()
        for a in range(0, len(counts), group):
            b = min(a + group, len(counts))
            nmax = int(counts[a:b].max())
            sub = {k: (v[a:b, :nmax].contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}
            xs0 = x0[a:b, :nmax].contiguous()
            rec = []

            def f(t, x, sub=sub, rec=rec):
                v = m(sub, x, t * torch.ones(x.shape[0]))
                if not rec:
                    rec.append(v.clone())
                rec.append(None)
                return v
            with torch.no_grad():
                xs = odeint(f, xs0, torch.linspace(0, 1, n_steps), method=method)
            sm = sub["q_mask"]
            finals.append(xs[-1][sm][:, 0]); mids.append(xs[n_steps // 2][sm][:, 0]); v0s.append(rec[0][sm][:, 0])
            nfe = len(rec) - 1
            print(method, f"events {a}:{b} nmax {nmax} {time.time()-t0:.0f}s", flush=True)
        out[method] = {"x_final": torch.cat(finals), "x_mid": torch.cat(mids), "v0": torch.cat(v0s), "nfe": nfe}
    assert out["euler"]["x_final"].numel() == int(mask.sum())
    torch.save(out, os.path.join(OUT, fname))


if __name__ == "__main__":
    torch.manual_seed(0)
    cases = sys.argv[1:] or ["taps", "dopri5", "config1", "multipart"]
    if "taps" in cases:
        taps_case("single_e", [12, 40, 24], 48, "sr_taps_single_e.pt")
        taps_case("multipart", [16, 80, 48, 32], 96, "sr_taps_multipart.pt")
    if "dopri5" in cases:
        dopri5_case("sr_dopri5_single_e.pt")
    if "config1" in cases:
        config1_case("sr_config1_single_e.pt")
    if "multipart" in cases:
        multipart_case("sr_traj_multipart.pt")
