"""Post-processing rows: CPU checks of the oracle against the reference's own transformation
classes (when /root/reference is present) and GPU parity of the device kernels (include/srhep_post.h) against the oracle."""
import sys

import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as po
from oracle import ref_import
from superresolutionhep_b200 import postprocess
from superresolutionhep_b200.default_configs import pflow_var_transform
from superresolutionhep_b200.synthetic import synthetic_events

TARGET = {"transformation": "logit_ratio", "f": 1.2, "alpha": 1.0e-6, "scale_mode": "standard", "mean": -1.1424768, "std": 3.616942}


def sr_batch(counts, seed=1):
    b = synthetic_events("single_e", len(counts), seed=seed, counts=np.array(counts))
    g = torch.Generator().manual_seed(seed)
    b["e_proxy_raw"] = (torch.rand(b["e_proxy"].shape, generator=g) * 5 + 0.01) * b["q_mask"].unsqueeze(-1)        # GeV
    return b


def test_stored_steps_match_reference_rule():
    ts, idx = postprocess.stored_steps(25, 4)
    assert idx == [0, 6, 12, 18] and ts == [0.0, 0.25, 0.5, 0.75]
    assert postprocess.stored_steps(25, 0) == ([], [])
    assert postprocess.stored_steps(25, -1) == ([], [])              # configs/multipart/inference_batch.yml:15


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")
def test_oracle_transforms_equal_reference_classes():
    if ref_import.REF_ROOT not in sys.path:
        sys.path.append(ref_import.REF_ROOT)
    from utility.target_transformation import TargetTransformation
    from utility.transformation import VarTransformation
    y, pr = torch.randn(1000) * 3, torch.rand(1000) * 4 + 0.01
    torch.testing.assert_close(po.target_inverse(TARGET, y, pr), TargetTransformation(dict(TARGET)).inverse(y, pr), rtol=1e-6, atol=1e-8)
    from oracle.pflow_oracle import var_forward
    for k, cfg in pflow_var_transform().items():
        x = torch.rand(100) * 50 + 1 if k != "eta" else torch.rand(100) * 5 - 2.5
        torch.testing.assert_close(var_forward(cfg, x), VarTransformation(dict(cfg)).forward(x), rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("n_ens,n_store", [(1, 0), (3, 4), (10, 2)])
def test_ensemble_unscale_matches_per_event_loop(n_ens, n_store):
    counts = [12, 132, 4, 64]
    batch = sr_batch(counts)
    n_steps = 9
    g = torch.Generator().manual_seed(3)
    comps = [torch.randn(n_steps, len(counts), batch["q_mask"].shape[1], 1, generator=g) * 2 for _ in range(n_ens)]
    ts, idx = postprocess.stored_steps(n_steps, n_store)
    ref = po.fill_high_tree(TARGET, batch, comps, ts, idx)
    mask = batch["q_mask"]
    keep = idx + [n_steps - 1]
    samples = torch.stack([c[keep][..., 0][:, mask] for c in comps], 0).cuda()                        # (E, S, T)
    nn_avg, e_avg, e_raw = postprocess.ensemble_unscale(samples, batch["e_proxy_raw"].reshape(mask.shape)[mask].cuda(), TARGET)
    cat = lambda k: torch.from_numpy(np.concatenate(ref[k]))
    tol = dict(rtol=1e-4, atol=1e-4)                                                                   # MeV
    torch.testing.assert_close(nn_avg[-1].cpu(), cat("raw_nn_pred"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(e_avg[-1].cpu(), cat("e_pred_avg_raw"), **tol)
    torch.testing.assert_close(e_raw[-1].cpu(), cat("e_pred_raw"), **tol)
    for j, t in enumerate(ts):
        torch.testing.assert_close(e_avg[j].cpu(), cat(f"e_pred_avg_raw_{t:.2f}"), **tol)
        torch.testing.assert_close(e_raw[j].cpu(), cat(f"e_pred_raw_{t:.2f}"), **tol)
        torch.testing.assert_close(nn_avg[j].cpu(), cat(f"raw_nn_pred_{t:.2f}"), rtol=1e-5, atol=1e-6)
    ev = postprocess.split_events(e_raw[-1], counts)
    assert [len(a) for a in ev] == counts and np.allclose(ev[1], ref["e_pred_raw"][1], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_sr_to_pflow_selection_matches_dataset_rule():
    rng = np.random.default_rng(0)
    counts = [300, 1, 0, 257, 64, 1000]
    e = [rng.exponential(3.0, n).astype(np.float32) for n in counts]                 # MeV: a good fraction below the 1 MeV cut
    e[1][:] = 0.5                                                                     # an event that loses every cell
    eta = [rng.uniform(-2.5, 2.5, n).astype(np.float32) for n in counts]
    phi = [rng.uniform(-np.pi, np.pi, n).astype(np.float32) for n in counts]
    lay = [rng.integers(0, 3, n).astype(np.int32) for n in counts]
    vt = pflow_var_transform()
    ref = po.pflow_cells_from_sr(e, eta, phi, lay, vt, 1.0)
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs)).to(dt).cuda()
    got = postprocess.sr_to_pflow(cat(e, torch.float32), cat(eta, torch.float32), cat(phi, torch.float32), cat(lay, torch.int32), counts, vt, 1.0)
    assert torch.equal(got["cell_mask"].cpu(), ref["cell_mask"])                      # which cells survive, per event, bit-exact
    for k in ("cell_e_raw", "cell_eta_raw", "cell_phi"):
        assert torch.equal(got[k].cpu(), ref[k]), k                                   # copied values, original order
    assert torch.equal(got["cell_layer"].cpu(), ref["cell_layer"])
    for k in ("cell_e", "cell_eta", "cell_cosphi", "cell_sinphi"):
        torch.testing.assert_close(got[k].cpu(), ref[k], rtol=1e-5, atol=1e-6, msg=k)


@pytest.mark.gpu
def test_ensemble_sample_end_to_end_shapes_and_mean():
    """Sampler -> ensemble -> unscale on the device, against the reference's loop fed with the same trajectories."""
    from superresolutionhep_b200 import FlowModel
    from superresolutionhep_b200.default_configs import flow_config
    from superresolutionhep_b200.synthetic import synthetic_state_dict
    m = FlowModel(flow_config("single_e"), precision="fp32")
    m.load_state_dict(synthetic_state_dict(m.dims, seed=2)); m.eval().cuda()
    counts = [24, 132, 8]
    batch = sr_batch(counts, seed=4)
    E, n_steps = 3, 5
    x0 = torch.randn(E, len(counts), batch["q_mask"].shape[1], 1, generator=torch.Generator().manual_seed(9))
    out = postprocess.ensemble_sample(m, batch, TARGET, n_ensemble=E, n_steps=n_steps, n_steps_to_store=2, method="euler", x0=x0)
    dev_batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    comps = [m.generate_samples(dev_batch, n_steps=n_steps, method="euler", ret_seq=True, x0=x0[e].cuda()).cpu() for e in range(E)]
    ts, idx = postprocess.stored_steps(n_steps, 2)
    ref = po.fill_high_tree(TARGET, batch, comps, ts, idx)
    assert out["stored_times"] == ts and out["counts"].tolist() == counts
    for k in ("e_pred_raw", "e_pred_avg_raw", "raw_nn_pred", f"e_pred_raw_{ts[1]:.2f}"):
        torch.testing.assert_close(out[k].cpu(), torch.from_numpy(np.concatenate(ref[k])), rtol=1e-4, atol=1e-4, msg=k)


@pytest.mark.gpu
def test_sr_to_pflow_pipeline_on_device(golden_dir):
    """BASELINE.json configs[4] end to end on the device: unscaled SR energies -> cell selection -> SAPF (real pf_hr weights),
    against the CPU pipeline (dataset rule + pflow oracle) on the same cells."""
    import os
    from oracle import pflow_oracle
    from superresolutionhep_b200.pflow import PflowLightning
    g = torch.load(os.path.join(golden_dir, "pflow_pf_hr.pt"))
    lm = PflowLightning({"pf_model": g["pf_model"], "var_transform": g["var_transform"]}, {}, inference=True)
    lm.load_state_dict({"net." + k: v for k, v in g["state_dict"].items()}, strict=True)
    lm.eval().cuda()
    rng = np.random.default_rng(5)
    counts = [400, 96, 640]
    e = [(rng.exponential(20.0, n) + 0.2).astype(np.float32) for n in counts]
    eta = [(rng.uniform(-2, 2) + rng.normal(0, 0.1, n)).astype(np.float32) for n in counts]
    phi = [(rng.uniform(-3, 3) + rng.normal(0, 0.1, n)).astype(np.float32) for n in counts]
    lay = [rng.integers(0, 3, n).astype(np.int32) for n in counts]
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs)).to(dt).cuda()
    batch = postprocess.sr_to_pflow(cat(e, torch.float32), cat(eta, torch.float32), cat(phi, torch.float32), cat(lay, torch.int32), counts, g["var_transform"], 1.0)
    logits, kin, inc = lm.net(batch)
    ref_batch = po.pflow_cells_from_sr(e, eta, phi, lay, g["var_transform"], 1.0)
    with torch.no_grad():
        lo, kin_r, inc_r, _ = pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], ref_batch)
    assert torch.equal(logits.argmax(-1).cpu(), lo.argmax(-1))
    torch.testing.assert_close(logits.cpu(), lo, rtol=1e-4, atol=1e-4 * float(lo.abs().max()))
    torch.testing.assert_close(kin.cpu(), kin_r, rtol=1e-4, atol=1e-4 * float(kin_r.abs().max()))
    torch.testing.assert_close(inc.cpu(), inc_r, rtol=1e-4, atol=1e-5)
