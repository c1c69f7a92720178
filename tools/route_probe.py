import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch, __graft_entry__ as ge, bench
pkg = ge.load_package(); dev = torch.device("cuda")
nq, dim = 1_000_000, 128
q = bench.synth_rows(nq, dim, 2002, dev)
cent = np.random.default_rng(0).standard_normal((4, dim)).astype(np.float32)
for i in range(4):
    torch.cuda.synchronize(); t = time.perf_counter()
    dest = pkg.route_queries(cent, q.data_ptr(), nq, slack=0.25)
    print(f"route_queries: {1e3 * (time.perf_counter() - t):.1f} ms", np.bincount(dest, minlength=4), flush=True)
