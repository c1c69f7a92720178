"""Scratch (ONE GPU): what the partitioned kernel variant costs when nothing is remote.  The ranks of a 2-way partition are
emulated on one device (raw-pointer attach), so every "remote" read is a local HBM read: the difference to the full index is
the partitioned kernel variant's own overhead (read ids, statistics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import __graft_entry__ as ge
import bench
n, nq, ef, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 2
dim = 128
pkg = ge.load_package()
dev = torch.device("cuda")
base = bench.synth_rows(n, dim, 1001, dev)
q = bench.synth_rows(nq, dim, 2002, dev)
torch.cuda.synchronize()
ix = pkg.Index.build_device(base.data_ptr(), n, dim, 16, 200)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
ref = torch.empty((nq, 10), dtype=torch.int32, device=dev)
def best(h, out):
    h.search_device(q.data_ptr(), nq, 10, ef, out.data_ptr())
    return min(h.search_device(q.data_ptr(), nq, 10, ef, out.data_ptr())["kernel_ms"] for _ in range(3))
t_full = best(ix, ref)
owner = torch.empty(n, dtype=torch.uint8, device=dev)
cent, sizes = ix.placement_fit(world, owner.data_ptr(), seed=1234, slack=0.05)
for mode in ("cluster", "round-robin"):
    parts = [ix.partition(r, world, 8, d_owner=owner.data_ptr() if mode == "cluster" else 0) for r in range(world)]
    ex = [p.partition_export() for p in parts]
    for r, p in enumerate(parts):
        for peer in range(world):
            if peer != r:
                p.partition_attach(peer, raw_ptrs=ex[peer][2])
    torch.cuda.synchronize()
    t_part = best(parts[0], ids)
    st = parts[0].search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr())
    tot = st["rows_hot"] + st["rows_local"] + st["rows_remote"] + st["rows_halo"]
    print(f"{mode}: full {t_full:.1f} ms, partition (all local HBM) {t_part:.1f} ms (+{100 * (t_part / t_full - 1):.1f}%), "
          f"identical {(ids == ref).all(1).float().mean().item():.4f}, hot {st['rows_hot'] / tot:.3f} own {st['rows_local'] / tot:.3f} other {st['rows_remote'] / tot:.3f}", flush=True)
    for p in parts:
        p.close()
