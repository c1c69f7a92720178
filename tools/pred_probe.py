import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, __graft_entry__ as ge, bench
pkg = ge.load_package(); dev = torch.device("cuda")
n, nq = 10_000_000, 200000
base = bench.synth_rows(n, 128, 1001, dev); q = bench.synth_rows(nq, 128, 2002, dev)
ix = pkg.Index.build_device(base.data_ptr(), n, 128, 16, 200)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
for ef in (16, 32, 64, 128, 256):
    st = ix.search_device(q.data_ptr(), nq, 10, ef, ids.data_ptr())
    print(f"ef={ef}: expansions/q {st['lists_l0'] / nq:.1f}, predictions {st['rows_hot'] / nq:.1f}, hits {st['rows_local'] / nq:.1f} -> accuracy {st['rows_local'] / max(1, st['rows_hot']):.3f}", flush=True)
