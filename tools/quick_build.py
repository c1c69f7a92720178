"""Scratch probe: GPU build of N rows, then ef sweep with recall vs torch brute force."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import __graft_entry__ as ge
import bench

n = int(sys.argv[1]); dim = int(sys.argv[2]); nq = int(sys.argv[3]); efc = int(sys.argv[4]) if len(sys.argv) > 4 else 200
ip = len(sys.argv) > 5 and sys.argv[5] == "ip"
efs = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [16, 32, 64, 128, 256]
pkg = ge.load_package()
dev = torch.device("cuda")
base = bench.synth_rows(n, dim, 1001, dev, normalize=ip)
q = bench.synth_rows(nq, dim, 2002, dev, normalize=ip)
torch.cuda.synchronize()
t = time.time()
ix = pkg.Index.build_device(base.data_ptr(), n, dim, 16, efc, ip=ip)
bs = ix.build_stats()
print(f"build n={n} dim={dim} efc={efc}: {time.time() - t:.1f}s wall, {bs['kernel_ms'] / 1e3:.1f}s device, "
      f"{n / bs['kernel_ms'] * 1e3 / 1e3:.1f}k inserts/s, distcomps/insert {bs['distcomps'] / n:.0f}, max_level {ix.max_level}", flush=True)
gt = bench.ground_truth(pkg, base, q[:5000].contiguous(), 10, ip, 0)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
dists = torch.empty((nq, 10), dtype=torch.float32, device=dev)
for ef in efs:
    ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())
    best = 1e9
    for _ in range(3):
        s = ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())
        best = min(best, s["kernel_ms"])
    rec = bench.recall_at_k(ids[:5000], gt)
    gbs = s["algorithmic_bytes"] / best / 1e6
    print(f"ef={ef:3d}: {best:8.2f} ms  {nq / best / 1e3:8.3f} MQPS  recall {rec:.4f}  distcomps/q {s['distcomps'] / nq:.0f}  "
          f"alg {gbs:.0f} GB/s ({gbs / 6550.1:.1%})  ovf {s['overflow_queries']}", flush=True)
