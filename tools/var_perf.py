"""Scratch: QPS of one library variant (SHN_LIB) on an n x dim GPU-built index."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import __graft_entry__ as ge
import bench
n, dim, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
efs = [int(x) for x in sys.argv[4].split(",")]
caps = [int(x) for x in sys.argv[5].split(",")]
pkg = ge.load_package()
dev = torch.device("cuda")
base = bench.synth_rows(n, dim, 1001, dev)
q = bench.synth_rows(nq, dim, 2002, dev)
ix = pkg.Index.build_device(base.data_ptr(), n, dim, 16, 200)
bs = ix.build_stats()
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
dists = torch.empty((nq, 10), dtype=torch.float32, device=dev)
out = [f"build {bs['kernel_ms'] / 1e3:.1f}s"]
for cap in caps:
    ix.set_option("visited_smem_entries", cap)
    for ef in efs:
        ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())
        best = min(ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr(), dists.data_ptr())["kernel_ms"] for _ in range(3))
        out.append(f"cap{cap}/ef{ef}: {nq / best / 1e3:.3f}M")
print(os.path.basename(os.environ.get("SHN_LIB", "default")), f"n={n} d={dim}:", "  ".join(out), flush=True)
