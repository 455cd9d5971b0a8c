"""CPU: the oracle restatement against the committed golden vectors (minted from the
reference's own modules by tests/golden/make_golden.py) and, when /root/reference is
mounted (build container), against the reference modules directly."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_import, sr_oracle
from oracle.odeint import odeint
from superresolutionhep_b200.config import SrDims
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict


def _setup(kind, seed):
    cfg = flow_config(kind)
    d = SrDims.from_config(cfg)
    sd = synthetic_state_dict(d, seed=seed)
    return cfg, d, sd, sr_oracle.derive_dims(cfg)


@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_oracle_matches_golden_taps(kind, golden_dir):
    g = torch.load(os.path.join(golden_dir, f"sr_taps_{kind}.pt"))
    cfg, d, sd, dims = _setup(kind, g["weight_seed"])
    batch = synthetic_events(kind, len(g["counts"]), seed=g["event_seed"], counts=np.array(g["counts"]), pad_to=g["pad_to"])
    x = synthetic_noise(batch, seed=g["noise_seed"])
    taps = {}
    with torch.no_grad():
        sr_oracle.flow_forward(sd, dims, batch, x, g["t"], taps=taps)
    for k, ref in g["taps"].items():
        torch.testing.assert_close(taps[k], ref, rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")


def test_oracle_matches_golden_config1_subset(golden_dir):
    """Events are independent under a fixed grid, so the first 6 of the 64 golden events,
    re-run alone (shorter padding), must reproduce their golden rows."""
    g = torch.load(os.path.join(golden_dir, "sr_config1_single_e.pt"))
    cfg, d, sd, dims = _setup("single_e", g["weight_seed"])
    full = synthetic_events("single_e", g["n_events"], seed=g["event_seed"])
    x0 = synthetic_noise(full, seed=g["noise_seed"])
    sel = 6
    n = full["q_mask"][:sel].sum(1)
    nmax = int(n.max())
    sub = {k: (v[:sel, :nmax] if torch.is_tensor(v) else v) for k, v in full.items()}
    with torch.no_grad():
        xs = sr_oracle.generate_samples(sd, dims, sub, x0[:sel, :nmax], n_steps=g["n_steps"], method="euler", ret_seq=True)
    m = sub["q_mask"]
    ref = g["euler"]["x_final"][:sel, :nmax]
    torch.testing.assert_close(xs[-1][m], ref[m], rtol=1e-4, atol=1e-4)
    refm = g["euler"]["x_mid"][:sel, :nmax]
    torch.testing.assert_close(xs[g["n_steps"] // 2][m], refm[m], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("method", ["euler", "midpoint"])
def test_oracle_matches_golden_multipart_subset(method, golden_dir):
    """The six shortest of the 16 golden multipart events (16 ... 320 cells), re-run alone: their packed golden rows
    (the golden is stored packed in entry order) must be reproduced by the oracle."""
    g = torch.load(os.path.join(golden_dir, "sr_traj_multipart.pt"))
    cfg, d, sd, dims = _setup("multipart", g["weight_seed"])
    counts = np.array(g["counts"])
    full = synthetic_events("multipart", len(counts), seed=g["event_seed"], counts=counts)
    x0 = synthetic_noise(full, seed=g["noise_seed"])
    a = len(counts) - 6
    nmax = int(counts[a:].max())
    sub = {k: (v[a:, :nmax].contiguous() if torch.is_tensor(v) else v) for k, v in full.items()}
    with torch.no_grad():
        xs = sr_oracle.generate_samples(sd, dims, sub, x0[a:, :nmax], n_steps=g["n_steps"], method=method, ret_seq=True)
    m = sub["q_mask"]
    off = int(counts[:a].sum())
    torch.testing.assert_close(xs[-1][m][:, 0], g[method]["x_final"][off:], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(xs[g["n_steps"] // 2][m][:, 0], g[method]["x_mid"][off:], rtol=1e-4, atol=1e-4)


def test_oracle_dopri5_matches_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "sr_dopri5_single_e.pt"))
    cfg, d, sd, dims = _setup("single_e", g["weight_seed"])
    batch = synthetic_events("single_e", g["n_events"], seed=g["event_seed"])
    x0 = synthetic_noise(batch, seed=g["noise_seed"])
    with torch.no_grad():
        xs = sr_oracle.generate_samples(sd, dims, batch, x0, n_steps=g["n_steps"], method="dopri5", ret_seq=True)
    torch.testing.assert_close(xs, g["x_seq"], rtol=1e-4, atol=1e-5)


def test_odeint_fixed_grid_orders():
    """Known-answer: dy/dt = -y, y(0)=1 -> e^{-1}; error orders 1 / 2 / 4; dopri5 within tol."""
    f = lambda t, y: -y
    y0 = torch.ones(3, dtype=torch.float64)
    exact = float(np.exp(-1.0))
    errs = {}
    for method in ("euler", "midpoint", "rk4"):
        e = []
        for n in (11, 21):
            sol = odeint(f, y0, torch.linspace(0, 1, n, dtype=torch.float64), method=method)
            assert sol.shape == (n, 3) and torch.equal(sol[0], y0)
            e.append(abs(float(sol[-1, 0]) - exact))
        errs[method] = np.log2(e[0] / e[1])
    assert 0.8 < errs["euler"] < 1.2 and 1.8 < errs["midpoint"] < 2.2 and 3.7 < errs["rk4"] < 4.3
    st = {}
    sol = odeint(f, y0.float(), torch.linspace(0, 1, 5), method="dopri5", stats=st)
    assert abs(float(sol[-1, 0]) - exact) < 1e-4 and st["nfe"] == 2 + 6 * (st["accepted"] + st["rejected"])


def test_padding_invariance_of_oracle():
    """real rows do not depend on padding or on x at padded slots."""
    cfg, d, sd, dims = _setup("single_e", 7)
    b1 = synthetic_events("single_e", 2, seed=5, counts=np.array([16, 24]))
    b2 = synthetic_events("single_e", 2, seed=5, counts=np.array([16, 24]), pad_to=96)
    x1 = synthetic_noise(b1, 1)
    x2 = torch.randn(b2["e_proxy"].shape)
    x2[:, :24] = x1
    t = torch.tensor([0.2, 0.7])
    with torch.no_grad():
        v1 = sr_oracle.flow_forward(sd, dims, b1, x1, t)
        v2 = sr_oracle.flow_forward(sd, dims, b2, x2, t)
    m = b1["q_mask"]
    torch.testing.assert_close(v1[m], v2[:, :24][m], rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted on this machine")
@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_oracle_vs_reference_modules(kind):
    cfg, d, sd, dims = _setup(kind, 3)
    m = ref_import.build_reference_flow_model(cfg, sd)
    assert set(m.state_dict().keys()) == set(sd.keys())
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == d.param_shapes()
    batch = synthetic_events(kind, 4, seed=9, counts=np.array([16, 64, 32, 48]))
    x = synthetic_noise(batch, 2)
    t = torch.tensor([0.0, 0.25, 0.5, 1.0])
    with torch.no_grad():
        torch.testing.assert_close(sr_oracle.flow_forward(sd, dims, batch, x, t), m(batch, x, t), rtol=0, atol=0)
