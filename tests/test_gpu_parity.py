"""GPU parity tests: the CUDA path (through the C ABI, via the FlowModel drop-in) against
the committed golden vectors (minted from the reference's own modules) and against the CPU
oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 path ``rtol 1e-4`` with ``atol = ATOL_EVAL * max|ref|``
for one network evaluation and ``atol = 1e-4 * max|ref|`` after a full ODE trajectory
(24-48 chained evaluations).  ``precision="fp32"`` runs on the tensor cores by default (every 16-bit
operand an fp16 (hi, lo) pair, products as hi.hi + hi.lo + lo.hi, fp32 accumulation: measured
rel-L2 2e-6 on the transformer output, <= 3e-5 * max|ref| on v) -> ATOL_EVAL = 5e-5; the CUDA-core
reference-order path (SRHEP_FP32_SIMT=1: rel-L2 6e-7, <= 1e-5 * max|ref| on v) is held to 2e-5 in
``test_fp32_cuda_core_path_matches_golden_and_tensor_core_path``; bf16 path ``rtol 1e-2`` with ``atol = 1e-2 * max|ref|``
(bare elementwise rtol 1e-2 is not met even by PyTorch's own bf16 autocast of
the reference).  Masks are passed through and must be bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import sr_oracle
from superresolutionhep_b200 import FlowModel, SupResLightning
from superresolutionhep_b200.config import SrDims
from superresolutionhep_b200.default_configs import flow_config, model_and_var_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_state_dict

pytestmark = pytest.mark.gpu

ATOL_EVAL = 5e-5          # x max|ref|, one evaluation, default (tensor-core) fp32 path
ATOL_EVAL_SIMT = 2e-5     # the CUDA-core fp32 path


def close(got, ref, rtol, atol_frac, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    atol = atol_frac * float(ref.abs().max())
    torch.testing.assert_close(got, ref, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def to_dev(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def make_model(kind, seed, precision="fp32"):
    cfg = flow_config(kind)
    m = FlowModel(cfg, precision=precision)
    sd = synthetic_state_dict(m.dims, seed=seed)
    m.load_state_dict(sd)
    return m.eval().cuda(), sd, sr_oracle.derive_dims(cfg)


def packed(t, mask):
    return t.reshape(mask.shape)[mask]


# ------------------------------------------------------------------------------------ golden
@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_fp32_cuda_core_path_matches_golden_and_tensor_core_path(kind, golden_dir, monkeypatch):
    """SRHEP_FP32_SIMT=1 (read by srhep_create): reference-order fp32 on the CUDA cores, held to the tighter bound; the
    default fp32 path (tensor cores, split operands) must agree with it far inside the north-star tolerance."""
    g = torch.load(os.path.join(golden_dir, f"sr_taps_{kind}.pt"))
    batch = synthetic_events(kind, len(g["counts"]), seed=g["event_seed"], counts=np.array(g["counts"]), pad_to=g["pad_to"])
    x = synthetic_noise(batch, seed=g["noise_seed"])
    mask = batch["q_mask"]
    names = ["feat_0", "layer_0", "transformer_out"]
    monkeypatch.setenv("SRHEP_FP32_SIMT", "1")
    m_simt, _, _ = make_model(kind, g["weight_seed"])
    taps_s, _ = m_simt.debug_taps(to_dev(batch), x.cuda(), g["t"].cuda(), names)        # the handle is created here, with the switch set
    monkeypatch.delenv("SRHEP_FP32_SIMT")
    m_tc, _, _ = make_model(kind, g["weight_seed"])
    taps_t, _ = m_tc.debug_taps(to_dev(batch), x.cuda(), g["t"].cuda(), names)
    assert m_simt.launch_count != m_tc.launch_count                                      # two different schedules really ran
    for n in names:
        close(taps_s[n], g["taps"][n][mask], 1e-4, ATOL_EVAL_SIMT, f"simt {n}")
        close(taps_t[n], taps_s[n], 1e-4, ATOL_EVAL, f"tensor-core vs cuda-core {n}")
    close(taps_s["v_t"].cpu()[mask], g["taps"]["v_t"][mask], 1e-4, ATOL_EVAL_SIMT, "simt v_t")
    close(taps_t["v_t"].cpu()[mask], taps_s["v_t"].cpu()[mask], 1e-4, ATOL_EVAL, "tensor-core vs cuda-core v_t")


@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_taps_match_golden(kind, golden_dir):
    g = torch.load(os.path.join(golden_dir, f"sr_taps_{kind}.pt"))
    m, sd, dims = make_model(kind, g["weight_seed"])
    batch = synthetic_events(kind, len(g["counts"]), seed=g["event_seed"], counts=np.array(g["counts"]), pad_to=g["pad_to"])
    x = synthetic_noise(batch, seed=g["noise_seed"])
    names = ["time_emb", "context", "feat_0", "layer_0", f"layer_{m.dims.layers - 1}", "transformer_out"]
    taps, ev = m.debug_taps(to_dev(batch), x.cuda(), g["t"].cuda(), names)
    mask = batch["q_mask"]
    for n in names:
        ref = g["taps"][n]
        if n in ("time_emb", "context"):
            close(taps[n], ref, 1e-4, ATOL_EVAL, n)
        else:
            close(taps[n], ref[mask], 1e-4, ATOL_EVAL, n)
    close(taps["v_t"].cpu()[mask], g["taps"]["v_t"][mask], 1e-4, ATOL_EVAL, "v_t")
    assert torch.equal(ev.mask.cpu(), mask)
    # padded slots of forward() are zero
    assert float(taps["v_t"].cpu()[~mask].abs().max()) == 0.0 if (~mask).any() else True


@pytest.mark.parametrize("method", ["euler", "midpoint"])
def test_config1_trajectory_matches_golden(method, golden_dir):
    """BASELINE.json configs[0]: 64 single-electron events, fixed noise seed, n_steps = 25."""
    g = torch.load(os.path.join(golden_dir, "sr_config1_single_e.pt"))
    m, sd, dims = make_model("single_e", g["weight_seed"])
    batch = synthetic_events("single_e", g["n_events"], seed=g["event_seed"])
    x0 = synthetic_noise(batch, seed=g["noise_seed"])
    xs = m.generate_samples(to_dev(batch), n_steps=g["n_steps"], method=method, ret_seq=True, x0=x0.cuda())
    assert xs.shape == (g["n_steps"],) + tuple(x0.shape)
    mask = batch["q_mask"]
    assert m.last_stats["nfe"] == g[method]["nfe"]
    close(xs[-1].cpu()[mask], g[method]["x_final"][mask], 1e-4, 1e-4, "x_final")
    close(xs[g["n_steps"] // 2].cpu()[mask], g[method]["x_mid"][mask], 1e-4, 1e-4, "x_mid")
    assert torch.equal(xs[0].cpu(), x0)
    # first velocity of the trajectory = forward at t = 0 on x0
    v0 = m(to_dev(batch), x0.cuda(), torch.zeros(g["n_events"]).cuda())
    close(v0.cpu()[mask], g[method]["v_evals"][0][mask], 1e-4, ATOL_EVAL, "v(t=0)")
    # ret_seq=False returns the last state
    xl = m.generate_samples(to_dev(batch), n_steps=g["n_steps"], method=method, ret_seq=False, x0=x0.cuda())
    assert torch.equal(xl.cpu()[mask], xs[-1].cpu()[mask])


@pytest.mark.parametrize("method", ["euler", "midpoint"])
def test_multipart_trajectory_matches_golden(method, golden_dir):
    """BASELINE.json configs[2] shapes, fp32 path: 16 multipart events (16 ... 3280 cells) as one batch against the
    reference's own trajectories (tests/golden/make_golden.py multipart)."""
    g = torch.load(os.path.join(golden_dir, "sr_traj_multipart.pt"))
    m, sd, dims = make_model("multipart", g["weight_seed"])
    batch = synthetic_events("multipart", len(g["counts"]), seed=g["event_seed"], counts=np.array(g["counts"]))
    x0 = synthetic_noise(batch, seed=g["noise_seed"])
    mask = batch["q_mask"]
    xs = m.generate_samples(to_dev(batch), n_steps=g["n_steps"], method=method, ret_seq=True, x0=x0.cuda()).cpu()
    assert m.last_stats["nfe"] == g[method]["nfe"]
    close(xs[-1][mask][:, 0], g[method]["x_final"], 1e-4, 1e-4, "x_final")
    close(xs[g["n_steps"] // 2][mask][:, 0], g[method]["x_mid"], 1e-4, 1e-4, "x_mid")
    v0 = m(to_dev(batch), x0.cuda(), torch.zeros(len(g["counts"])).cuda()).cpu()
    close(v0[mask][:, 0], g[method]["v0"], 1e-4, ATOL_EVAL, "v(t=0)")


def test_dopri5_matches_golden(golden_dir):
    """Adaptive solver: both sides integrate the same ODE to atol = rtol = 1e-4; the step
    sequences differ (the reference's error norm includes padded slots), so the comparison
    is at solver tolerance, not at fp32 rounding."""
    g = torch.load(os.path.join(golden_dir, "sr_dopri5_single_e.pt"))
    m, sd, dims = make_model("single_e", g["weight_seed"])
    batch = synthetic_events("single_e", g["n_events"], seed=g["event_seed"])
    x0 = synthetic_noise(batch, seed=g["noise_seed"])
    xs = m.generate_samples(to_dev(batch), n_steps=g["n_steps"], method="dopri5", ret_seq=True, x0=x0.cuda())
    mask = batch["q_mask"]
    st = m.last_stats
    assert st["nfe"] == 2 + 6 * (st["accepted"] + st["rejected"])
    for j in range(g["n_steps"]):
        close(xs[j].cpu()[mask], g["x_seq"][j][mask], 2e-3, 2e-3, f"x_seq[{j}]")


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_device_resident_dopri5_equals_host_driven_loop(precision):
    """The adaptive loop as one conditional CUDA graph (controller, accept / reject, dense output on the device; no host
    round trip per step) against the same arithmetic driven from the host (use_graph = False, one scalar read back per
    attempted step): same accepted / rejected step sequence, same outputs up to the controller's libm-vs-device `pow`
    and FMA contraction of the interpolation coefficients.  Multi-pass bindings and ret_seq = False included."""
    m, sd, dims = make_model("single_e", 19, precision)
    counts = np.array([124, 8, 132, 256, 64, 300])
    batch = synthetic_events("single_e", len(counts), seed=31, counts=counts)
    x0 = synthetic_noise(batch, seed=32)
    mask = batch["q_mask"]
    db = to_dev(batch)
    dev = m.generate_samples(db, n_steps=6, method="dopri5", ret_seq=True, x0=x0.cuda()).cpu()
    st_dev = dict(m.last_stats)
    again = m.generate_samples(db, n_steps=6, method="dopri5", ret_seq=True, x0=x0.cuda()).cpu()       # cached graph, second launch
    assert torch.equal(again[:, mask], dev[:, mask]) or precision == "fp32"                            # (double atomics order the norm sums freely: last-bit dt differences are allowed)
    last = m.generate_samples(db, n_steps=6, method="dopri5", ret_seq=False, x0=x0.cuda()).cpu()
    close(last[mask], dev[-1][mask], 1e-5, 1e-5, "ret_seq=False")
    m.pass_tokens = 300                                                                                # several passes inside the graph body
    cut = m.generate_samples(db, n_steps=6, method="dopri5", ret_seq=True, x0=x0.cuda()).cpu()
    close(cut[:, mask], dev[:, mask], 1e-5, 1e-5, "multi-pass")
    m.pass_tokens = 0
    m.use_graph = False
    host = m.generate_samples(db, n_steps=6, method="dopri5", ret_seq=True, x0=x0.cuda()).cpu()
    st_host = dict(m.last_stats)
    print(f"[{precision}] device {st_dev} host {st_host}")
    assert st_dev == st_host and st_dev["nfe"] == 2 + 6 * (st_dev["accepted"] + st_dev["rejected"])
    assert torch.equal(dev[0], x0)
    close(dev[:, mask], host[:, mask], 1e-5, 1e-5, "device vs host dopri5")


# ------------------------------------------------------------------------------------ oracle
@pytest.mark.parametrize("kind,counts", [
    ("single_e", [4, 128, 132, 36, 260]),
    ("multipart", [16, 0, 304, 48, 16, 0]),         # events with zero cells (multipart min is 0)
])
def test_forward_matches_oracle(kind, counts):
    m, sd, dims = make_model(kind, 21)
    batch = synthetic_events(kind, len(counts), seed=3, counts=np.array(counts))
    x = synthetic_noise(batch, seed=4)
    t = torch.linspace(0.0, 1.0, len(counts))
    v = m(to_dev(batch), x.cuda(), t.cuda()).cpu()
    keep = [i for i, c in enumerate(counts) if c > 0]          # the oracle (like the reference) NaNs on n = 0
    sub = {k: (val[keep] if torch.is_tensor(val) else val) for k, val in batch.items()}
    with torch.no_grad():
        ref = sr_oracle.flow_forward(sd, dims, sub, x[keep], t[keep])
    mask = sub["q_mask"]
    close(v[keep][mask], ref[mask], 1e-4, ATOL_EVAL, "v")
    assert torch.isfinite(v).all()


def test_rk4_and_single_point_grid_match_oracle():
    m, sd, dims = make_model("single_e", 5)
    batch = synthetic_events("single_e", 3, seed=8, counts=np.array([20, 64, 8]))
    x0 = synthetic_noise(batch, seed=9)
    with torch.no_grad():
        ref = sr_oracle.generate_samples(sd, dims, batch, x0, n_steps=4, method="rk4", ret_seq=True)
    xs = m.generate_samples(to_dev(batch), n_steps=4, method="rk4", ret_seq=True, x0=x0.cuda())
    mask = batch["q_mask"]
    close(xs.cpu()[:, mask], ref[:, mask], 1e-4, 1e-4, "rk4")
    assert m.last_stats["nfe"] == 12
    one = m.generate_samples(to_dev(batch), n_steps=1, method="euler", ret_seq=True, x0=x0.cuda())
    assert one.shape[0] == 1 and torch.equal(one[0].cpu(), x0)


def test_passes_graphs_and_padding_do_not_change_results():
    """Events are independent: cutting the batch into passes, replaying a captured graph
    instead of launching directly, and extra padding must all give identical real rows."""
    m, sd, dims = make_model("single_e", 13)
    counts = np.array([40, 8, 132, 64, 4, 96, 256, 12])
    b1 = synthetic_events("single_e", len(counts), seed=2, counts=counts)
    b2 = synthetic_events("single_e", len(counts), seed=2, counts=counts, pad_to=300)
    x1 = synthetic_noise(b1, seed=6)
    x2 = torch.zeros(b2["e_proxy"].shape)
    x2[:, : x1.shape[1]] = x1
    mask1, mask2 = b1["q_mask"], b2["q_mask"]
    ref = m.generate_samples(to_dev(b1), n_steps=5, method="midpoint", ret_seq=True, x0=x1.cuda()).cpu()[:, mask1]
    m.pass_tokens = 100
    a = m.generate_samples(to_dev(b1), n_steps=5, method="midpoint", ret_seq=True, x0=x1.cuda()).cpu()[:, mask1]
    m.use_graph = False
    b = m.generate_samples(to_dev(b1), n_steps=5, method="midpoint", ret_seq=True, x0=x1.cuda()).cpu()[:, mask1]
    m.pass_tokens = 0
    c = m.generate_samples(to_dev(b2), n_steps=5, method="midpoint", ret_seq=True, x0=x2.cuda()).cpu()[:, mask2]
    assert torch.equal(ref, a) and torch.equal(ref, b) and torch.equal(ref, c)


def test_lightning_shim_loads_prefixed_checkpoint(tmp_path):
    """inference.py:74-83: SupResLightning(...).load_state_dict(torch.load(ckpt)['state_dict'])."""
    cfg = model_and_var_config("multipart")
    d = SrDims.from_config(cfg["flow_model"])
    sd = synthetic_state_dict(d, seed=17)
    path = tmp_path / "epoch=0.ckpt"
    torch.save({"state_dict": {"net." + k: v for k, v in sd.items()}, "epoch": 0}, path)
    lm = SupResLightning(cfg, {}, precision="fp32")
    lm.load_state_dict(torch.load(path, map_location="cpu", weights_only=True)["state_dict"])
    lm.eval().cuda()
    batch = synthetic_events("multipart", 2, seed=1, counts=np.array([32, 16]))
    x = synthetic_noise(batch, seed=2)
    t = torch.tensor([0.3, 0.6])
    v = lm.net(to_dev(batch), x.cuda(), t.cuda()).cpu()
    with torch.no_grad():
        ref = sr_oracle.flow_forward(sd, sr_oracle.derive_dims(cfg["flow_model"]), batch, x, t)
    close(v[batch["q_mask"]], ref[batch["q_mask"]], 1e-4, ATOL_EVAL, "v")


def test_full_size_linearity_of_sharding():
    """BASELINE.json config 2 shape (bounded: 512 events): the two halves of a batch sampled
    separately reproduce the whole batch bit-exactly (entry-range sharding)."""
    m, sd, dims = make_model("single_e", 29)
    batch = synthetic_events("single_e", 512, seed=1234)
    x0 = synthetic_noise(batch, seed=0)
    whole = m.generate_samples(to_dev(batch), n_steps=3, method="euler", x0=x0.cuda()).cpu()
    mask = batch["q_mask"]
    for sl in (slice(0, 256), slice(256, 512)):
        sub = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in batch.items()}
        part = m.generate_samples(to_dev(sub), n_steps=3, method="euler", x0=x0[sl].cuda()).cpu()
        assert torch.equal(part[mask[sl]], whole[sl][mask[sl]])
    assert torch.isfinite(whole).all()


def test_cost_balanced_shards_concatenate_to_whole():
    """entry ranges planned by cost, sampled independently, concatenated in entry order
    = the whole batch, bit for bit (fixed-grid methods, fp32 path)."""
    from superresolutionhep_b200 import sharding
    m, sd, dims = make_model("multipart", 13)
    counts = np.array([16, 640, 32, 48, 16, 320, 64, 1008, 16])
    batch = synthetic_events("multipart", len(counts), seed=21, counts=counts)
    x0 = synthetic_noise(batch, seed=22)
    mask = batch["q_mask"]
    whole = m.generate_samples(to_dev(batch), n_steps=4, method="euler", x0=x0.cuda()).cpu()
    # world = 1 goes through the same code path without a process group
    via = sharding.sample_sharded(m, batch, n_steps=4, method="euler", x0=x0).cpu()
    assert torch.equal(via[mask], whole[mask])
    parts = []
    for a, b in sharding.plan_entry_ranges(counts, 3):
        sub = sharding.shard_batch(batch, a, b)
        xs = m.generate_samples(to_dev(sub), n_steps=4, method="euler", x0=x0[a:b, : sub["q_mask"].shape[1]].cuda()).cpu()
        parts.append(xs[..., 0][sub["q_mask"]])
    assert torch.equal(torch.cat(parts), whole[..., 0][mask])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shared_time_path_equals_per_event_time_path(precision):
    """A sampling pass evaluates every event at the same t: the timestep embedding is computed for one event and
    copied, and the time columns of the adaLN contraction are folded into its bias (srhep.cu: enqueue_eval).
    ``forward`` takes a time per event and uses the general path.  One Euler step from the sampler must reproduce
    ``x0 + dt * forward(x0, t0)``: the two paths differ only in fp32 summation order."""
    m, _, _ = make_model("single_e", 7, precision=precision)
    counts = np.array([124, 256, 4, 300, 804])
    batch = synthetic_events("single_e", len(counts), seed=11, counts=counts)
    x0 = synthetic_noise(batch, seed=3)
    db = to_dev(batch)
    mask = batch["q_mask"]
    n_steps = 5
    xs = m.generate_samples(db, n_steps=n_steps, method="euler", ret_seq=True, x0=x0.cuda())
    dt = 1.0 / (n_steps - 1)
    v = m(db, x0.cuda(), torch.zeros(len(counts), device="cuda"))
    want = packed(x0, mask) + dt * packed(v.cpu(), mask)
    got = packed(xs[1].cpu(), mask)
    tol = ATOL_EVAL if precision == "fp32" else 2e-3      # 16-bit operands: a last-bit difference of the adaLN rows moves a rounding boundary now and then
    close(got, want, tol, tol, f"shared-t vs per-event-t ({precision})")
