"""The N>1 host logic on CPU: world_size 2 and 3, gloo backend — sharding as the reference deals queries to compute
nodes, the single all-gather of per-rank top-k lists back into global order, rolling recall."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, k, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    par = ge.load_package().parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full_ids = rng.integers(0, 1 << 30, size=(n_total, k), dtype=np.int64).astype(np.int32)
    full_d = rng.random((n_total, k)).astype(np.float32)
    gt = full_ids.copy()
    gt[::3] = -5  # a third of the queries miss everything
    slots = par.shard_slots(n_total, rank, world)
    assert len(slots) == par.shard_size(n_total, rank, world)
    ids = torch.from_numpy(full_ids[slots])
    d = torch.from_numpy(full_d[slots])
    all_ids, all_d = par.allgather_results(ids, d, n_total, rank, world, dist)
    ok = bool((all_ids.numpy() == full_ids).all() and (all_d.numpy().view(np.uint32) == full_d.view(np.uint32)).all())
    rec = par.local_recall(ids, torch.from_numpy(gt[slots]))
    roll = par.rolling_recall(rec, len(slots), n_total, dist)
    want = par.local_recall(torch.from_numpy(full_ids), torch.from_numpy(gt))
    ok = ok and abs(roll - want) < 1e-12
    with open(os.path.join(out_dir, f"r{rank}"), "w") as f:
        f.write("ok" if ok else f"bad roll={roll} want={want}")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1001), (3, 64), (2, 1), (2, 0)])
def test_shard_and_gather(tmp_path, world, n_total):
    port = 29500 + (os.getpid() + world * 7 + n_total) % 2000
    mp.spawn(_worker, args=(world, port, n_total, 10, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "ok"


def test_shard_slots_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    par = ge.load_package().parallel
    for n in (0, 1, 7, 100):
        for world in (1, 2, 3, 8):
            parts = [par.shard_slots(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert [len(p) for p in parts] == [par.shard_size(n, r, world) for r in range(world)]


def _route_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    par = ge.load_package().parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 500 + 37 * rank
    rows = torch.from_numpy(rng.random((n, 6)).astype(np.float32))
    dest = rng.integers(0, world, n)
    if rank == 0:
        dest[:] = world - 1  # everything leaves rank 0
    ex = par.RoutedExchange(dest, world, dist, "cpu")
    got = ex.forward(rows)
    assert got.shape[0] == sum(ex.recv_list)
    res = torch.stack([got.sum(1), got[:, 0] * 2 + rank * 0], 1)  # "search": a function of the row only
    back = ex.backward(res)
    ok = bool(torch.allclose(back[:, 0], rows.sum(1)) and torch.allclose(back[:, 1], rows[:, 0] * 2))
    with open(os.path.join(out_dir, f"r{rank}"), "w") as f:
        f.write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_routed_exchange_round_trip(tmp_path, world):
    port = 31500 + (os.getpid() + world * 11) % 2000
    mp.spawn(_route_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "ok"
