import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, __graft_entry__ as ge, bench
pkg = ge.load_package(); dev = torch.device("cuda")
n, nq, ef = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
base = bench.synth_rows(n, 128, 1001, dev); q = bench.synth_rows(nq, 128, 2002, dev)
ix = pkg.Index.build_device(base.data_ptr(), n, 128, 16, 200)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
torch.cuda.profiler.start()
st = ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr())
torch.cuda.profiler.stop()
print(st["kernel_ms"], st["algorithmic_bytes"] / 1e9)
