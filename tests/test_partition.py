"""Partitioned index: the graph split over `world` GPUs with a replicated hot set.  With one GPU the ranks
are emulated in one process on one device (each rank's share is a separate allocation, "peer" pointers are attached
raw) — the addressing, the renumbering and the counters are what a multi-GPU run executes; results must be bit-identical
to the unpartitioned index."""
import numpy as np
import pytest

import datagen

pytestmark = pytest.mark.gpu


def make_parts(full, world, ratio):
    parts = [full.partition(r, world, ratio) for r in range(world)]
    exports = [p.partition_export() for p in parts]
    for r, p in enumerate(parts):
        for peer in range(world):
            if peer != r:
                p.partition_attach(peer, raw_ptrs=exports[peer][2])
    return parts


@pytest.mark.parametrize("world,ratio,ip", [(2, 5, False), (3, 10, False), (4, 0, True), (8, 5, False)])
def test_partitioned_search_is_bit_identical(pkg, world, ratio, ip):
    n, dim = 20000, 32
    base, queries = datagen.base_and_queries(n, 600, dim, normalize=ip)
    with pkg.Index.build(base, 16, 100, ip=ip) as full:
        full.count_visits(True)
        full.search(queries[300:], 10, 64)          # warm-up pass picks the hot set (compute_node.cc:116-131)
        ref = full.search(queries[:300], 10, 64)
        parts = make_parts(full, world, ratio)
        full.count_visits(False)
    try:
        totals = np.zeros(3)
        for r, p in enumerate(parts):
            assert p.n == n
            ids, dists, st = p.search(queries[:300], 10, 64)
            assert (ids == ref[0]).all() and (dists.view(np.uint32) == ref[1].view(np.uint32)).all(), f"rank {r}"
            for key in ("distcomps", "visited_nodes", "visited_nodes_l0", "visited_neighborlists"):
                assert st[key] == ref[2][key]
            # every level-0 row read is classified exactly once
            assert st["rows_hot"] + st["rows_local"] + st["rows_remote"] in (st["visited_nodes_l0"], st["visited_nodes_l0"] - 300)
            totals += [st["rows_hot"], st["rows_local"], st["rows_remote"]]
        hot, local, remote = totals / totals.sum()
        if ratio:
            assert hot > ratio / 100, "the hot set must serve more than its share of the reads"
        assert abs(remote / (local + remote) - (world - 1) / world) < 0.05  # round-robin placement of the cold rows
    finally:
        for p in parts:
            p.close()


def test_partition_state_errors(pkg):
    base, queries = datagen.base_and_queries(3000, 10, 16)
    with pkg.Index.build(base, 8, 40) as full:
        p0 = full.partition(0, 2, 5)
        with pytest.raises(pkg.ShnError) as e:   # peer 1 not attached yet
            p0.search(queries, 5, 20)
        assert e.value.code == -5
        with pytest.raises(pkg.ShnError):        # a share cannot be stored as a dump
            p0.to_dumps(1)
        with pytest.raises(pkg.ShnError):
            p0.partition(0, 2, 5)
        with pytest.raises(pkg.ShnError):
            full.partition(2, 2, 5)
        with pytest.raises(pkg.ShnError):
            p0.partition_attach(0, raw_ptrs=(1, 1))
        p0.close()


@pytest.mark.parametrize("world,ip", [(4, False), (2, True)])
def test_placement_by_cluster_and_routing(pkg, world, ip):
    """Nodes stored on the GPU of their nearest centroid, queries run on the GPU of theirs: same results, and most
    level-0 reads become local (the point of the reference's Placement + QueryRouter)."""
    import torch
    n, dim = 30000, 32
    base, queries = datagen.base_and_queries(n, 800, dim, normalize=ip)
    with pkg.Index.build(base, 16, 100, ip=ip) as full:
        ref = full.search(queries, 10, 64)
        owner = torch.empty(n, dtype=torch.uint8, device="cuda")
        cent, sizes = full.placement_fit(world, owner.data_ptr(), slack=0.05)
        assert sizes.sum() == n and sizes.max() <= 1.06 * n / world + 1 and cent.shape == (world, dim)
        assert (torch.bincount(owner.long(), minlength=world).cpu().numpy() == sizes).all()
        clustered = [full.partition(r, world, 0, d_owner=owner.data_ptr()) for r in range(world)]
        scattered = [full.partition(r, world, 0) for r in range(world)]
    for group in (clustered, scattered):
        ex = [p.partition_export() for p in group]
        for r, p in enumerate(group):
            for peer in range(world):
                if peer != r:
                    p.partition_attach(peer, raw_ptrs=ex[peer][2])
    try:
        q_dev = torch.from_numpy(queries).cuda()
        dest = pkg.route_queries(cent, q_dev.data_ptr(), len(queries), ip=ip, slack=0.25)
        assert dest.max() < world and np.bincount(dest, minlength=world).max() <= 1.25 * len(queries) / world + 1
        remote = {}
        for name, group in (("clustered", clustered), ("scattered", scattered)):
            rem = tot = 0
            for r, p in enumerate(group):
                mine = np.nonzero(dest == r)[0]
                if len(mine) == 0:
                    continue
                ids, dists, st = p.search(queries[mine], 10, 64)
                assert (ids == ref[0][mine]).all() and (dists.view(np.uint32) == ref[1][mine].view(np.uint32)).all()
                rem += st["rows_remote"]; tot += st["rows_hot"] + st["rows_local"] + st["rows_remote"]
            remote[name] = rem / tot
        print("remote fraction of level-0 reads:", remote)
        assert remote["clustered"] < 0.6 * remote["scattered"]
    finally:
        for p in clustered + scattered:
            p.close()
