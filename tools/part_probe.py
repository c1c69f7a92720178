"""Scratch: partitioned search with both shares in ONE process (raw peer pointers) on 2 GPUs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
n, dim, nq = int(sys.argv[1]), 128, 500000
parts = []
for g in range(2):
    dev = torch.device("cuda", g)
    torch.cuda.set_device(g)
    base = bench.synth_rows(n, dim, 1001, dev)
    full = pkg.Index.build_device(base.data_ptr(), n, dim, 16, 200, gpu=g)
    if len(sys.argv) > 2:
        warm = bench.synth_rows(100000, dim, 7, dev)
        tmp = torch.empty((100000, 10), dtype=torch.int32, device=dev)
        full.count_visits(True)
        full.search_device(warm.data_ptr(), 100000, 10, 64, tmp.data_ptr())
    parts.append(full.partition(g, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 0))
    if g == 0:
        q = bench.synth_rows(nq, dim, 2002, dev)
        ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
        ref = full.search_device(q.data_ptr(), nq, 10, 64, ids.data_ptr())
        ref = full.search_device(q.data_ptr(), nq, 10, 64, ids.data_ptr())
        print(f"full index on GPU0: {nq / ref['kernel_ms'] / 1e3:.3f} MQPS", flush=True)
        ref_ids = ids.clone()
    full.close(); del base
ex = [p.partition_export() for p in parts]
parts[0].partition_attach(1, raw_ptrs=ex[1][2]); parts[1].partition_attach(0, raw_ptrs=ex[0][2])
torch.cuda.set_device(0)
for _ in range(2):
    st = parts[0].search_device(q.data_ptr(), nq, 10, 64, ids.data_ptr())
tot = st["rows_hot"] + st["rows_local"] + st["rows_remote"]
print(f"partitioned (raw peer ptrs): {nq / st['kernel_ms'] / 1e3:.3f} MQPS, remote {st['rows_remote'] / tot:.3f}, hot {st['rows_hot'] / tot:.3f}, "
      f"nvlink in {st['rows_remote'] * 512 / st['kernel_ms'] / 1e6:.0f} GB/s, identical ids: {bool((ids == ref_ids).all())}", flush=True)
