"""Scratch perf probe (not the bench): reference-built index of N rows, GPU search at several ef."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import __graft_entry__ as ge
import datagen, shine_ref, hnsw_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
pkg = ge.load_package()
base, queries = datagen.base_and_queries(n, nq, dim)
t = time.time()
dumps, st, secs = shine_ref.build(base, m=16, efc=200, threads=min(32, os.cpu_count()), coroutines=4)
print(f"ref build n={n} dim={dim}: {secs:.1f}s with {min(32, os.cpu_count())} threads", flush=True)
gt = datagen.bruteforce(base, queries[:1000], 10)
ix = pkg.Index.from_dumps(dumps, dim, 16)
q = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
dists = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
for wps in (0, 1024, 2048, 4096):
    ix.set_option("visited_smem_entries", wps)
    for ef in (16, 64, 200, 256):
        ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())
        best = 1e9
        for _ in range(3):
            s = ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())
            best = min(best, s["kernel_ms"])
        rec = datagen.recall(ids[:1000].cpu().numpy().view(np.uint32), gt)
        gbs = s["algorithmic_bytes"] / best / 1e6
        print(f"wps={wps:2d} ef={ef:3d}: {best:8.2f} ms  {nq / best / 1e3:8.3f} MQPS  recall {rec:.3f}  "
              f"distcomps/q {s['distcomps'] / nq:.0f}  alg {gbs:.0f} GB/s ({gbs / 6550.1:.1%} of measured HBM)  ovf {s['overflow_queries']}", flush=True)
