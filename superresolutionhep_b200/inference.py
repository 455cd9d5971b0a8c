"""Drop-in for the reference's SR inference driver ``inference.py`` (boundary only).

Same command line and YAML styles::

    python -m superresolutionhep_b200.inference -i cfg.yml [-p highest|high|medium] [-bm] [-estart A] [-estop B]

and the same control flow as ``Inference`` (inference.py:39-160): read ``config_path_mv`` / ``config_path_t`` /
``checkpoint_path`` from the ``model`` block, derive the grid points to store (inference.py:54-69), load the
Lightning checkpoint's ``state_dict`` into ``SupResLightning`` (:74-83), loop the batches, draw ``n_ensemble``
samples with ``generate_samples(..., ret_seq=True)`` (:145-148), average and unscale them (:152-287), collect the
``Low_Tree`` / ``High_Tree`` / ``Particle_Tree`` branches and write them (:291-310).

What is different, and why:
* the ensemble members of a batch run as ONE batch and the ensemble mean + ``TargetTransformation.inverse`` run in one
  device kernel (postprocess.py) instead of the per-event Python loop -- same branch names and values;
* the ROOT reader (``dataset.SupResDataset``: uproot + DGL) and writer (uproot + awkward) are used when those packages
  and the reference's ``dataset.py`` are importable; otherwise batches must come from ``Inference.run_batches`` (any
  iterable of ``collate_graphs_plus`` dicts) and the trees are written as ``<pred_path>.npz`` with one object array per
  branch, keyed ``"<Tree>/<branch>"``.  Neither package is installable in the build image, so only the fall-back is tested;
* keys that the shipped single_e YAMLs do not define but the reference reads (``store_ensemble_components``,
  ``store_energy_incidence``, ``max_particles``: SURVEY.md Appendix D) default to False / False / 4 instead of raising;
* ``-p/--precision`` selects the arithmetic of the sm_100a path: ``highest`` = fp32-grade arithmetic on the tensor cores (every operand a
  pair of fp16 planes, fp32 accumulation; rel. error 2e-6 against the fp32 reference), ``high`` / ``medium`` = fp16
  tcgen05 operands with fp32 accumulation (the reference passes the same string to ``torch.set_float32_matmul_precision``, inference.py:346,374).
"""
from __future__ import annotations

import argparse
import os
import time
from pathlib import Path
from typing import Dict, Iterable, Optional

import numpy as np
import torch
import yaml

from . import postprocess
from .lightning import SupResLightning

HIGH_PASSTHROUGH = (("eta_raw", "eta_raw", 1.0), ("phi", "phi", 1.0), ("layer", "layer", 1.0), ("e_truth_raw", "e_truth_raw", 1e3),
                    ("e_proxy", "e_proxy", 1.0), ("e_proxy_raw", "e_proxy_raw", 1e3), ("raw_nn_cond", "e_proxy", 1.0), ("raw_nn_target", "target", 1.0))
LOW_PASSTHROUGH = (("eta_raw", "low_eta_raw", 1.0), ("phi", "low_phi", 1.0), ("layer", "low_layer", 1.0), ("e_meas_raw", "low_e_meas_raw", 1e3))
PARTICLE_KEYS = ("particle_pt", "particle_eta", "particle_phi", "particle_e", "particle_pdgid", "particle_dep_e")


class Inference:
    def __init__(self, inf_cfg: dict, device: Optional[torch.device] = None):
        self.inf_cfg = inf_cfg
        self.config_path_mv = inf_cfg["model"]["config_path_mv"]
        with open(self.config_path_mv) as fp:
            self.config_mv = yaml.safe_load(fp)
        self.config_path_t = inf_cfg["model"]["config_path_t"]
        with open(self.config_path_t) as fp:
            self.config_t = yaml.safe_load(fp)
        if device is None and not torch.cuda.is_available():
            raise RuntimeError("superresolutionhep_b200.inference needs a CUDA (sm_100a) device -- there is no CPU path")
        self.device = device or torch.device("cuda")
        self.load_model()
        self.n_steps = self.inf_cfg["model"]["n_steps"]
        self.ts_to_store, self.ts_to_store_idx = postprocess.stored_steps(self.n_steps, self.inf_cfg["model"]["n_steps_to_store"])   # inference.py:54-69
        self.target_cfg = self.config_mv["target_transform"]                                                                    # inference.py:71
        self.method = self.inf_cfg["model"].get("method", "dopri5")                 # the reference never passes `method`: torchdiffeq's dopri5

    def load_model(self):
        """inference.py:74-83."""
        self.lightning_model = SupResLightning(self.config_mv, self.config_t)
        checkpoint = torch.load(self.inf_cfg["model"]["checkpoint_path"], map_location=torch.device("cpu"), weights_only=True)
        self.lightning_model.load_state_dict(checkpoint["state_dict"])
        torch.set_grad_enabled(False)
        self.lightning_model.eval()
        self.lightning_model.to(self.device)

    # ------------------------------------------------------------------ input
    def get_dataloader(self, inf_dict: dict):
        """inference.py:86-93 -- needs the reference's ``dataset.py`` (uproot, awkward, dgl) on ``sys.path``."""
        try:
            from dataset import SupResDataset, collate_graphs_plus        # the reference's own reader
            from torch.utils.data import DataLoader
        except Exception as e:                                            # noqa: BLE001
            raise RuntimeError("reading ROOT files needs the reference's dataset.py with uproot, awkward and dgl installed "
                               f"({type(e).__name__}: {e}); feed collate_graphs_plus-style batches to Inference.run_batches instead") from e
        ds = SupResDataset(inf_dict["truth_path"], reduce_ds=inf_dict["n_events"], entry_start=inf_dict["entry_start"], config_mv=self.config_mv,
                           make_low_graph=True, make_particle_graph=True, one_event_train=self.config_t["one_event_train"],
                           one_event_idx=self.config_t["one_event_idx"])
        return DataLoader(ds, batch_size=inf_dict["batch_size"], num_workers=inf_dict["num_workers"], shuffle=False, collate_fn=collate_graphs_plus)

    # ------------------------------------------------------------------ branches
    def prep_dicts(self, inf_dict: dict):
        """inference.py:96-131."""
        self.low_dict_to_zip = {k: [] for k in ("eta_raw", "phi", "layer", "e_meas_raw")}
        self.high_dict_to_zip = {k: [] for k in ("eta_raw", "phi", "layer", "e_proxy", "e_truth_raw", "e_proxy_raw", "e_pred_raw", "e_pred_avg_raw",
                                                 "raw_nn_cond", "raw_nn_target", "raw_nn_pred")}
        for t in self.ts_to_store:
            for k in (f"e_pred_raw_{t:.2f}", f"e_pred_avg_raw_{t:.2f}", f"raw_nn_pred_{t:.2f}"):
                self.high_dict_to_zip[k] = []
        n_ens = inf_dict.get("n_ensemble", 1)
        self.store_components = n_ens > 1 and bool(inf_dict.get("store_ensemble_components", inf_dict.get("save_ensemble_components", False)))
        if self.store_components:
            for i in range(n_ens):
                self.high_dict_to_zip[f"e_pred_raw_comp_{i}"] = []
                self.high_dict_to_zip[f"raw_nn_pred_comp_{i}"] = []
                for t in self.ts_to_store:
                    self.high_dict_to_zip[f"e_pred_raw_{t:.2f}_comp_{i}"] = []
                    self.high_dict_to_zip[f"raw_nn_pred_{t:.2f}_comp_{i}"] = []
        self.particle_dict_to_zip = {k: [] for k in PARTICLE_KEYS}
        self.store_incidence = bool(inf_dict.get("store_energy_incidence", False))
        self.max_particles = int(inf_dict.get("max_particles", 4))
        if self.store_incidence:
            for i in range(self.max_particles):
                self.low_dict_to_zip[f"e_part_{i}"] = []
                self.high_dict_to_zip[f"e_part_{i}"] = []

    def _append_packed(self, dst: dict, key: str, packed: torch.Tensor, counts):
        if key in dst:
            dst[key].extend(postprocess.split_events(packed, counts))

    def fill_the_dicts2write(self, batch: Dict[str, torch.Tensor], res: dict, n_ens: int):
        """inference.py:163-287 with the sampler-dependent branches already reduced on the device (``res``)."""
        counts = res["counts"]
        hd = self.high_dict_to_zip
        mask = batch["q_mask"].bool()
        for name in ("e_pred_raw", "e_pred_avg_raw", "raw_nn_pred"):
            self._append_packed(hd, name, res[name], counts)
            for t in self.ts_to_store:
                self._append_packed(hd, f"{name}_{t:.2f}", res[f"{name}_{t:.2f}"], counts)
        if self.store_components:
            samples = res["samples"]                                          # (E, S, T); S = stored grid points + final
            proxy = batch["e_proxy_raw"].to(samples.device).reshape(mask.shape)[mask.to(samples.device)]
            for i in range(n_ens):
                _, _, e_i = postprocess.ensemble_unscale(samples[i:i + 1], proxy, self.target_cfg)
                self._append_packed(hd, f"e_pred_raw_comp_{i}", e_i[-1], counts)
                self._append_packed(hd, f"raw_nn_pred_comp_{i}", samples[i, -1], counts)
                for j, t in enumerate(self.ts_to_store):
                    self._append_packed(hd, f"e_pred_raw_{t:.2f}_comp_{i}", e_i[j], counts)
                    self._append_packed(hd, f"raw_nn_pred_{t:.2f}_comp_{i}", samples[i, j], counts)
        # pass-through branches (present when the batch comes from collate_graphs_plus)
        for out, key, unit in HIGH_PASSTHROUGH:
            if key in batch:
                v = batch[key].reshape(mask.shape)[mask.to(batch[key].device)]
                self._append_packed(hd, out, v * unit if unit != 1.0 else v, counts)
        if "low_q_mask" in batch:
            lm = batch["low_q_mask"].bool()
            lcounts = lm.sum(1).cpu().numpy()
            for out, key, unit in LOW_PASSTHROUGH:
                if key in batch:
                    v = batch[key].reshape(lm.shape)[lm.to(batch[key].device)]
                    self._append_packed(self.low_dict_to_zip, out, v * unit if unit != 1.0 else v, lcounts)
        if self.store_incidence:                                               # inference.py:266-273: per-particle energy incidence, zero rows up to max_particles
            if "low_e_part_0" not in batch or "high_e_part_0" not in batch or "particle_pt" not in batch:
                raise KeyError("store_energy_incidence needs low_e_part_i / high_e_part_i / particle_pt in the batch (collate_graphs_plus)")
            npy = lambda v: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
            sq = lambda a: a.squeeze(-1) if a.ndim and a.shape[-1] == 1 else a
            for bs_i in range(mask.shape[0]):
                n_part = len(batch["particle_pt"][bs_i])
                for pi in range(self.max_particles):
                    for tree, pre in ((self.low_dict_to_zip, "low"), (hd, "high")):
                        src = npy(batch[f"{pre}_e_part_{pi}"][bs_i]) if pi < n_part else np.zeros_like(npy(batch[f"{pre}_e_part_0"][bs_i]))
                        tree[f"e_part_{pi}"].append(sq(src))
        for k in PARTICLE_KEYS:
            if k in batch:
                self.particle_dict_to_zip[k].extend(np.asarray(p.detach().cpu().numpy() if torch.is_tensor(p) else p) for p in batch[k])

    # ------------------------------------------------------------------ driver
    def run_batches(self, batches: Iterable[Dict[str, torch.Tensor]], inf_dict: dict):
        """The loop of ``run_pred`` (inference.py:139-157) over any iterable of collate-style batch dicts."""
        self.prep_dicts(inf_dict)
        n_ens = int(inf_dict.get("n_ensemble", 1))
        net = self.lightning_model.net
        n_events = 0
        for batch in batches:
            batch = {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in batch.items()}
            res = postprocess.ensemble_sample(net, batch, self.target_cfg, n_ensemble=n_ens, n_steps=inf_dict.get("n_steps", self.n_steps),
                                              n_steps_to_store=self.inf_cfg["model"]["n_steps_to_store"], method=inf_dict.get("method", self.method))
            self.fill_the_dicts2write(batch, res, n_ens)
            n_events += int(batch["q_mask"].shape[0])
        return n_events

    def run_pred(self, inf_dict: dict):
        n = self.run_batches(self.get_dataloader(inf_dict), inf_dict)
        self.write_trees(inf_dict["pred_path"])
        return n

    def write_trees(self, pred_path: str) -> str:
        """inference.py:291-310: ROOT through uproot + awkward when installed, else one ``.npz`` with ``"<Tree>/<branch>"`` object arrays."""
        trees = {"Low_Tree": self.low_dict_to_zip, "High_Tree": self.high_dict_to_zip, "Particle_Tree": self.particle_dict_to_zip}
        try:
            import awkward as ak                                          # noqa: F401
            import uproot
            with uproot.recreate(pred_path) as file:
                for name, d in trees.items():
                    file[name] = {"": ak.zip({k: ak.Array(v) for k, v in d.items() if len(v)})}
            print(f"\nPredictions saved to {pred_path}")
            return pred_path
        except ImportError:
            out = pred_path[:-5] + ".npz" if pred_path.endswith(".root") else pred_path + ".npz"
            flat = {}
            for name, d in trees.items():
                for k, v in d.items():
                    arr = np.empty(len(v), dtype=object)
                    for i, a in enumerate(v):
                        arr[i] = np.asarray(a)
                    flat[f"{name}/{k}"] = arr
            np.savez(out, **flat)
            print(f"\nuproot/awkward not installed: predictions saved to {out}")
            return out

    def get_output_path(self, inf_dict: dict) -> str:
        """inference.py:313-325."""
        outputdir = os.path.join(os.path.dirname(self.config_path_mv), "inference")
        if inf_dict.get("dir_flag") is not None:
            outputdir = os.path.join(outputdir, inf_dict["dir_flag"])
        Path(outputdir).mkdir(parents=True, exist_ok=True)
        return os.path.join(outputdir, "{}_pred.root".format("_".join(inf_dict["truth_path"].split(".root")[0].split("/")[-1:])))


def build_parser() -> argparse.ArgumentParser:
    """inference.py:328-334, flag for flag."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--inference_path", "-i", type=str, required=True)
    ap.add_argument("--precision", "-p", type=str, required=False, default="highest")
    ap.add_argument("--batch_mode", "-bm", action="store_true")
    ap.add_argument("--entry_start", "-estart", type=int, required=False, default=0)
    ap.add_argument("--entry_stop", "-estop", type=int, required=False, default=None)
    return ap


def expand_inf_dicts(inference_cfg: dict, args) -> list:
    """The two YAML styles (inference.py:341-390): ``inf_dict`` for ``--batch_mode`` jobs over an entry range, ``items`` otherwise."""
    common = dict(num_workers=inference_cfg["num_workers"], batch_size=inference_cfg["batch_size"], n_steps=inference_cfg["model"]["n_steps"],
                  n_steps_to_store=inference_cfg["model"]["n_steps_to_store"], max_particles=inference_cfg.get("max_particles", 4))
    if args.batch_mode:
        if "items" in inference_cfg:
            raise ValueError("wrong config style for batch mode")
        if args.entry_stop is None:
            raise ValueError("entry_stop is required for batch mode")
        d = dict(inference_cfg["inf_dict"], **common)
        d["entry_start"], d["n_events"] = args.entry_start, args.entry_stop - args.entry_start
        d["_suffix"] = f"_{args.entry_start}_{args.entry_stop}"
        return [d]
    if "items" not in inference_cfg:
        raise ValueError("wrong config style for not batch mode")
    return [dict(item, gpu=inference_cfg["gpu"], **common) for item in inference_cfg["items"] if item.get("run_pred")]


def main(argv=None):
    args = build_parser().parse_args(argv)
    with open(args.inference_path) as fp:
        inference_cfg = yaml.safe_load(fp)
    inf_dicts = expand_inf_dicts(inference_cfg, args)
    torch.set_float32_matmul_precision(args.precision)                      # read by FlowModel: highest -> fp32 kernels, else fp16 tcgen05 operands
    if str(inference_cfg.get("gpu", -1)) not in ("-1", "None"):
        os.environ["CUDA_VISIBLE_DEVICES"] = str(inference_cfg["gpu"])
    inf_obj = Inference(inference_cfg)
    for inf_dict in inf_dicts:
        print("Running predictions on {}".format(inf_dict["truth_path"]))
        pred_path = inf_obj.get_output_path(inf_dict)
        inf_dict["pred_path"] = pred_path.replace(".root", inf_dict.pop("_suffix", "") + ".root")
        t1 = time.time()
        n = inf_obj.run_pred(inf_dict)
        print(f"Prediction time: {time.time() - t1:.2f} s ({n} events)")


if __name__ == "__main__":
    main()
