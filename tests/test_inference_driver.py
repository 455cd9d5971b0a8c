"""The inference.py drop-in: command line, the two YAML styles, the stored-grid rule and -- on the GPU --
one tiny end-to-end run from a Lightning-style checkpoint file to the written trees."""
import os

import numpy as np
import pytest
import torch
import yaml

from superresolutionhep_b200 import inference as inf
from superresolutionhep_b200.config import SrDims
from superresolutionhep_b200.default_configs import flow_config, model_and_var_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_state_dict

TARGET = {"transformation": "logit_ratio", "f": 1.2, "alpha": 1.0e-6, "scale_mode": "standard", "mean": -1.1424768, "std": 3.616942}


def test_cli_flags_match_the_reference():
    a = inf.build_parser().parse_args(["-i", "cfg.yml"])
    assert (a.inference_path, a.precision, a.batch_mode, a.entry_start, a.entry_stop) == ("cfg.yml", "highest", False, 0, None)
    a = inf.build_parser().parse_args(["--inference_path", "c.yml", "-p", "medium", "-bm", "-estart", "100", "-estop", "200"])
    assert (a.precision, a.batch_mode, a.entry_start, a.entry_stop) == ("medium", True, 100, 200)


def test_both_yaml_styles():
    base = {"gpu": 1, "num_workers": 2, "batch_size": 500, "model": {"n_steps": 25, "n_steps_to_store": 5}}
    items = dict(base, items=[{"run_pred": True, "truth_path": "a/test.root", "n_ensemble": 10}, {"run_pred": False, "truth_path": "b.root"}])
    args = inf.build_parser().parse_args(["-i", "x"])
    d = inf.expand_inf_dicts(items, args)
    assert len(d) == 1 and d[0]["n_steps"] == 25 and d[0]["batch_size"] == 500 and d[0]["max_particles"] == 4      # missing key defaults
    with pytest.raises(ValueError, match="not batch mode"):
        inf.expand_inf_dicts(dict(base, inf_dict={}), args)
    bm = inf.build_parser().parse_args(["-i", "x", "-bm", "-estart", "300", "-estop", "400"])
    d = inf.expand_inf_dicts(dict(base, max_particles=4, inf_dict={"truth_path": "t/train.root", "n_ensemble": 10}), bm)
    assert d[0]["entry_start"] == 300 and d[0]["n_events"] == 100 and d[0]["_suffix"] == "_300_400"
    with pytest.raises(ValueError, match="wrong config style for batch mode"):
        inf.expand_inf_dicts(items, bm)
    with pytest.raises(ValueError, match="entry_stop"):
        inf.expand_inf_dicts(dict(base, inf_dict={}), inf.build_parser().parse_args(["-i", "x", "-bm"]))


def _write_model_dir(tmp_path, n_store=2, n_steps=5):
    mv = model_and_var_config("single_e")
    mv["target_transform"] = TARGET
    (tmp_path / "config_mv.yml").write_text(yaml.safe_dump(mv))
    (tmp_path / "config_t.yml").write_text(yaml.safe_dump({"one_event_train": False, "one_event_idx": 0}))
    sd = synthetic_state_dict(SrDims.from_config(flow_config("single_e")), seed=11)
    torch.save({"state_dict": {"net." + k: v for k, v in sd.items()}, "epoch": 1}, tmp_path / "last.ckpt")
    return {"gpu": -1, "num_workers": 0, "batch_size": 3,
            "model": {"config_path_mv": str(tmp_path / "config_mv.yml"), "config_path_t": str(tmp_path / "config_t.yml"),
                      "checkpoint_path": str(tmp_path / "last.ckpt"), "n_steps": n_steps, "n_steps_to_store": n_store, "method": "euler"}}


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU path"):
        inf.Inference(_write_model_dir(tmp_path))


@pytest.mark.gpu
def test_end_to_end_from_checkpoint_to_trees(tmp_path):
    cfg = _write_model_dir(tmp_path)
    torch.set_float32_matmul_precision("highest")                       # -p highest -> fp32 kernels
    obj = inf.Inference(cfg)
    assert obj.ts_to_store_idx == [0, 2] and obj.lightning_model.net._device().type == "cuda"
    batches = []
    for seed, counts in ((1, [24, 132, 8]), (2, [64, 4])):
        b = synthetic_events("single_e", len(counts), seed=seed, counts=np.array(counts))
        b["e_proxy_raw"] = torch.rand(b["e_proxy"].shape) * 4 + 0.05
        b["eta_raw"] = b["eta"] * 2.988
        batches.append(b)
    inf_dict = {"n_ensemble": 3, "save_ensemble_components": True, "truth_path": "x/test.root", "n_steps": 5}
    assert obj.run_batches(batches, inf_dict) == 5
    hd = obj.high_dict_to_zip
    assert [len(a) for a in hd["e_pred_raw"]] == [24, 132, 8, 64, 4]
    for k in ("e_pred_avg_raw", "raw_nn_pred", "e_pred_raw_0.00", "e_pred_raw_0.50", "e_pred_raw_comp_2", "raw_nn_pred_0.50_comp_1", "eta_raw", "e_proxy_raw"):
        assert len(hd[k]) == 5 and hd[k][1].shape == (132,), k
    # t = 0 is the noise: unscaled energies at the first stored grid point are finite and positive-bounded by f * proxy
    assert np.all(np.isfinite(hd["e_pred_raw"][1])) and np.all(hd["e_pred_raw"][1] <= 1.2 * 1e3 * batches[0]["e_proxy_raw"][1, :132, 0].numpy() * (1 + 1e-5) + 1e-3)
    # ensemble mean of the members equals the stored mean
    comps = np.stack([hd[f"e_pred_raw_comp_{i}"][1] for i in range(3)])
    np.testing.assert_allclose(comps.mean(0), hd["e_pred_raw"][1], rtol=1e-5, atol=1e-4)
    out = obj.write_trees(obj.get_output_path(inf_dict))
    assert out.endswith("test_pred.npz") and os.path.isfile(out)
    z = np.load(out, allow_pickle=True)
    assert "High_Tree/e_pred_raw" in z.files and len(z["High_Tree/e_pred_raw"]) == 5 and z["High_Tree/e_pred_raw"][3].shape == (64,)
    with pytest.raises(RuntimeError, match="dataset.py"):
        obj.get_dataloader({"truth_path": "x.root", "n_events": 1, "entry_start": 0, "batch_size": 1, "num_workers": 0})


def test_energy_incidence_branches_are_filled_and_zero_padded():
    """inference.py:266-273: ``e_part_i`` of Low_Tree / High_Tree come from ``low_e_part_i`` / ``high_e_part_i`` for the event's
    particles and are zero arrays (shaped like particle 0's) up to ``max_particles``."""
    obj = inf.Inference.__new__(inf.Inference)
    obj.ts_to_store, obj.ts_to_store_idx, obj.target_cfg = [], [], TARGET
    obj.prep_dicts({"n_ensemble": 1, "store_energy_incidence": True, "max_particles": 3})
    counts = np.array([8, 4])
    b = synthetic_events("single_e", 2, seed=3, counts=counts)
    b["low_q_mask"] = torch.tensor([[True, True], [True, False]])
    b["particle_pt"] = [torch.rand(2), torch.rand(1)]
    for pi in range(2):
        b[f"low_e_part_{pi}"] = [torch.rand(2, 1) + pi, torch.rand(1, 1) + pi]
        b[f"high_e_part_{pi}"] = [torch.rand(8, 1) + pi, torch.rand(4, 1) + pi]
    T = int(counts.sum())
    res = {"counts": counts, **{k: torch.zeros(T) for k in ("e_pred_raw", "e_pred_avg_raw", "raw_nn_pred")}}
    obj.fill_the_dicts2write(b, res, 1)
    hd, ld = obj.high_dict_to_zip, obj.low_dict_to_zip
    assert [a.shape for a in hd["e_part_0"]] == [(8,), (4,)] and [a.shape for a in ld["e_part_1"]] == [(2,), (1,)]
    np.testing.assert_array_equal(hd["e_part_1"][0], b["high_e_part_1"][0][:, 0].numpy())
    assert np.all(hd["e_part_1"][1] == 0) and hd["e_part_1"][1].shape == (4,)         # event 1 has one particle: index 1 is padding
    assert np.all(hd["e_part_2"][0] == 0) and np.all(ld["e_part_2"][1] == 0)
    with pytest.raises(KeyError, match="store_energy_incidence"):
        obj.fill_the_dicts2write({k: v for k, v in b.items() if not k.startswith("low_e_part")}, res, 1)
