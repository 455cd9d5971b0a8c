"""Event sharding across the GPUs of one box.

Events never interact on the fixed-grid sampling path (attention is within an event, the context is
per event, the update is per cell), so the reference scales by independent jobs over entry ranges
(``inference.py -bm -estart A -estop B``, inference.py:341-367) and concatenates their outputs
(pflow/dataset_pf.py:29-30).  The in-process equivalent keeps that contract:

* ``plan_entry_ranges``: contiguous entry ranges, one per rank, cut where the cumulative COST (not the
  event count) crosses k/world of the total -- the cost of an event with n cells is a*n + b*n^2
  (the same n^2 budget idea as the reference's utility/sampler.py:24-45);
* ``shard_batch``: the rows ``[A, B)`` of a padded collate dict, trimmed to the shard's own Nmax;
* ``gather_packed``: the ONE collective of the path -- packed outputs (real cells only) of every rank
  to ``dst`` (NCCL over NVLink on the GPU box, gloo in the CPU tests); returns the concatenation in
  entry order, which is what the concatenated job outputs of the reference are.
No collective runs inside the sampling loop.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

SR_COST = (5088448.0, 6144.0, 3215360.0)        # F(n) = a n + b n^2 + c per network evaluation
PFLOW_COST = (214000.0, 768.0, 60000.0)         # ~214 kFLOP per cell + 768 n^2 per event


def event_cost(counts: Sequence[int], coef: Tuple[float, float, float] = SR_COST) -> np.ndarray:
    n = np.asarray(counts, dtype=np.float64)
    return coef[0] * n + coef[1] * n * n + coef[2]


def plan_entry_ranges(counts: Sequence[int], world: int, coef: Tuple[float, float, float] = SR_COST) -> List[Tuple[int, int]]:
    """``world`` contiguous ranges ``[A_r, B_r)`` covering ``[0, len(counts))`` in order, with
    cumulative cost as close as a prefix cut allows to r/world of the total.  Ranges may be empty
    when there are fewer events than ranks."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n_ev = len(counts)
    if n_ev == 0:
        return [(0, 0)] * world
    cum = np.concatenate([[0.0], np.cumsum(event_cost(counts, coef))])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        j = int(np.searchsorted(cum, target, side="left"))
        if j > 0 and (j > n_ev or abs(cum[j - 1] - target) <= abs(cum[min(j, n_ev)] - target)):
            j -= 1                                           # nearer of the two neighbouring prefix sums
        cuts.append(min(max(j, cuts[-1]), n_ev))
    cuts.append(n_ev)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_batch(batch: Dict[str, Optional[torch.Tensor]], start: int, stop: int, mask_key: str = "q_mask") -> Dict[str, Optional[torch.Tensor]]:
    """Events ``[start, stop)`` of a padded batch dict, padded only up to the shard's longest event."""
    mask = batch[mask_key][start:stop]
    nmax = int(mask.sum(1).max()) if stop > start and mask.numel() else 0
    nmax = max(nmax, 1)
    full = batch[mask_key].shape[1]
    out: Dict[str, Optional[torch.Tensor]] = {}
    for k, v in batch.items():
        if not torch.is_tensor(v):
            out[k] = v
            continue
        s = v[start:stop]
        if s.dim() >= 2 and s.shape[1] == full:
            s = s[:, :nmax]
            if s.dim() >= 3 and s.shape[2] == full:                  # (B, N, N) masks
                s = s[:, :, :nmax]
        out[k] = s.contiguous()
    return out


def gather_packed(local: torch.Tensor, counts_local: Sequence[int], dst: int = 0, group=None) -> Optional[Tuple[torch.Tensor, np.ndarray]]:
    """Gathers per-rank packed outputs ``(..., T_r)`` (cells of the rank's events, entry order) to ``dst``.

    Returns ``(packed (..., sum T_r), counts (all events))`` on ``dst`` and ``None`` elsewhere.  One
    size exchange (all_gather of two ints) plus one gather of the payload padded to the largest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local, np.asarray(counts_local, dtype=np.int64)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local.device
    lead = tuple(local.shape[:-1])
    sizes = torch.tensor([local.shape[-1], len(counts_local)], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    t_max = max(int(s[0]) for s in all_sizes)
    e_max = max(int(s[1]) for s in all_sizes)
    pay = torch.zeros(*lead, max(t_max, 1), dtype=local.dtype, device=dev)
    pay[..., : local.shape[-1]] = local
    cnt = torch.zeros(max(e_max, 1), dtype=torch.int64, device=dev)
    cnt[: len(counts_local)] = torch.as_tensor(np.asarray(counts_local, dtype=np.int64), device=dev)
    pays = [torch.empty_like(pay) for _ in range(world)] if rank == dst else None
    cnts = [torch.empty_like(cnt) for _ in range(world)] if rank == dst else None
    dist.gather(pay, pays, dst=dst, group=group)
    dist.gather(cnt, cnts, dst=dst, group=group)
    if rank != dst:
        return None
    packed = torch.cat([p[..., : int(s[0])] for p, s in zip(pays, all_sizes)], dim=-1)
    counts = np.concatenate([c[: int(s[1])].cpu().numpy() for c, s in zip(cnts, all_sizes)])
    return packed, counts


def unpack_to_padded(packed: torch.Tensor, counts: Sequence[int], nmax: Optional[int] = None, fill: float = 0.0) -> torch.Tensor:
    """``(..., T)`` packed cells -> ``(..., B, Nmax, 1)`` (the reference's padded output layout)."""
    counts = np.asarray(counts, dtype=np.int64)
    B = len(counts)
    nmax = int(max(counts.max() if B else 0, 1)) if nmax is None else nmax
    mask = torch.arange(nmax, device=packed.device).unsqueeze(0) < torch.as_tensor(counts, device=packed.device).unsqueeze(1)
    out = packed.new_full((*packed.shape[:-1], B, nmax), fill)
    out[..., mask] = packed
    return out.unsqueeze(-1)


def sample_sharded(model, batch: Dict[str, Optional[torch.Tensor]], n_steps: Optional[int] = None, method: str = "euler", ret_seq: bool = False,
                   x0: Optional[torch.Tensor] = None, dst: int = 0, group=None):
    """``FlowModel.generate_samples`` over the whole batch with the events split across the ranks of
    ``group``: every rank passes the SAME full batch (and, for reproducibility, the same ``x0``), samples
    its own entry range and rank ``dst`` gets the result in the reference layout ``([n_steps,] B, Nmax, 1)``
    (``None`` on the other ranks).  dopri5 couples the events of a batch through its error norm
    (SURVEY.md 8e), so only the fixed-grid methods are accepted when world > 1."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if world > 1 and method == "dopri5":
        raise ValueError("dopri5 couples the events of a batch through its global error norm; shard with euler / midpoint / rk4")
    counts = batch["q_mask"].sum(1).cpu().numpy()
    a, b = plan_entry_ranges(counts, world)[rank]
    sub = shard_batch(batch, a, b)
    dev = next(model.parameters()).device
    sub = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sub.items()}
    nmax_sub = sub["q_mask"].shape[1]
    x0_sub = None if x0 is None else x0[a:b, :nmax_sub].to(dev)
    if b > a:
        xs = model.generate_samples(sub, n_steps=n_steps, method=method, ret_seq=ret_seq, x0=x0_sub)
        packed = xs[..., 0][..., sub["q_mask"].bool()]                  # ([n_steps,] T_r)
    else:
        n_out = (n_steps if n_steps is not None else model.n_steps) if ret_seq else None
        packed = torch.zeros((n_out, 0) if ret_seq else (0,), dtype=torch.float32, device=dev)
    got = gather_packed(packed, counts[a:b], dst=dst, group=group)
    if got is None:
        return None
    full, all_counts = got
    return unpack_to_padded(full, all_counts, nmax=batch["q_mask"].shape[1])
