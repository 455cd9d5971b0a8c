"""CPU: the multi-GPU host logic -- entry-range planning, shard extraction and the one
gather of packed outputs, the latter over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superresolutionhep_b200 import sharding
from superresolutionhep_b200.synthetic import cell_counts, synthetic_events


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_plan_covers_in_order_and_balances_cost(kind, world):
    counts = cell_counts(kind, 4096, np.random.default_rng(0))
    ranges = sharding.plan_entry_ranges(counts, world)
    assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == len(counts)
    for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
        assert b == c and a <= b
    cost = sharding.event_cost(counts)
    per = np.array([cost[a:b].sum() for a, b in ranges])
    assert per.max() / per.mean() < 1.01                    # within 1 % of perfect balance at 4096 events
    by_count = np.array([cost[i * len(counts) // world:(i + 1) * len(counts) // world].sum() for i in range(world)])
    assert per.max() <= by_count.max() * 1.001              # never worse than the equal-count split


def test_plan_edge_cases():
    assert sharding.plan_entry_ranges([], 4) == [(0, 0)] * 4
    r = sharding.plan_entry_ranges([100, 100], 4)
    assert r[0][0] == 0 and r[-1][1] == 2 and sum(b - a for a, b in r) == 2
    assert sharding.plan_entry_ranges([3280, 16, 16, 16], 2) == [(0, 1), (1, 4)]          # one long event outweighs three short ones
    with pytest.raises(ValueError):
        sharding.plan_entry_ranges([1], 0)


def test_shard_batch_trims_padding():
    batch = synthetic_events("multipart", 6, seed=1, counts=np.array([16, 640, 32, 48, 16, 320]))
    sub = sharding.shard_batch(batch, 2, 5)
    assert sub["q_mask"].shape == (3, 48) and sub["eta"].shape == (3, 48, 1) and sub["edge_mask"] is None
    assert torch.equal(sub["e_proxy"][1, :48], batch["e_proxy"][3, :48])
    assert sub["q_mask"].sum(1).tolist() == [32, 48, 16]


def test_unpack_to_padded_roundtrip():
    counts = [3, 0, 5]
    packed = torch.arange(8, dtype=torch.float32)
    out = sharding.unpack_to_padded(packed, counts)
    assert out.shape == (3, 5, 1)
    assert out[0, :3, 0].tolist() == [0, 1, 2] and out[2, :, 0].tolist() == [3, 4, 5, 6, 7] and float(out[1].abs().sum()) == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, counts, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ranges = sharding.plan_entry_ranges(counts, world)
        a, b = ranges[rank]
        cu = np.concatenate([[0], np.cumsum(counts)])
        full = torch.arange(3 * cu[-1], dtype=torch.float32).reshape(3, -1)          # (n_keep = 3, T): value encodes (step, global cell)
        local = full[:, cu[a]:cu[b]].contiguous()
        got = sharding.gather_packed(local, counts[a:b], dst=0)
        if rank == 0:
            packed, all_counts = got
            ret["ok"] = bool(torch.equal(packed, full)) and all_counts.tolist() == list(counts)
        else:
            ret["none_%d" % rank] = got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[16, 640, 32, 48, 16, 320, 64], [5], [7, 9]])
def test_gather_packed_world2_gloo(counts):
    """Two ranks with different numbers of events and cells (incl. an empty shard): rank 0 receives the
    concatenation in entry order."""
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gather_worker, args=(world, _free_port(), list(counts), ret), nprocs=world, join=True)
        assert ret["ok"] is True and ret["none_1"] is True


def test_packed_events_pack_unpack_roundtrip():
    """PackedEvents (flow_model.py) finds the positions of the real cells once and packs / unpacks with index_select /
    index_copy_: same result as boolean-mask indexing of the padded (B, Nmax, 1) tensors (dataset.py:341-349 layout)."""
    import torch
    from superresolutionhep_b200.flow_model import PackedEvents
    from superresolutionhep_b200.synthetic import synthetic_events
    counts = np.array([124, 4, 256, 300])
    batch = synthetic_events("single_e", len(counts), seed=3, counts=counts)
    ev = PackedEvents(batch, torch.device("cpu"))
    m = batch["q_mask"]
    assert ev.n_cells == int(counts.sum()) and ev.cu_host.tolist() == [0] + np.cumsum(counts).tolist()
    for key in ("eta", "cosphi", "sinphi", "e_proxy"):
        assert torch.equal(getattr(ev, key), batch[key].reshape(m.shape)[m].float())
    assert torch.equal(ev.layer, batch["layer"].reshape(m.shape)[m].to(torch.int32))
    x = torch.randn(batch["e_proxy"].shape)
    p = ev.pack(x)
    assert torch.equal(p, x.reshape(m.shape)[m])
    u = ev.unpack(p)
    assert u.shape == x.shape and torch.equal(u[..., 0][m], p) and float(u[..., 0][~m].abs().sum()) == 0.0
    seq = torch.randn(3, ev.n_cells)
    us = ev.unpack(seq)
    assert us.shape == (3,) + tuple(x.shape)
    for j in range(3):
        assert torch.equal(us[j][..., 0][m], seq[j])
    uf = ev.unpack(p, fill=torch.full(x.shape, 7.0))
    assert float(uf[..., 0][~m].min()) == 7.0 and torch.equal(uf[..., 0][m], p)
