import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, __graft_entry__ as ge, bench
pkg = ge.load_package(); dev = torch.device("cuda")
n, nq, dim = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
base = bench.synth_rows(n, dim, 1001, dev); q = bench.synth_rows(nq, dim, 2002, dev)
out = {}
for mode in ("simt", "tc"):
    os.environ["SHN_BRUTEFORCE"] = mode
    ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        if rep == 1: torch.cuda.profiler.start()
        pkg.bruteforce_topk_device(base.data_ptr(), n, q.data_ptr(), nq, dim, 10, ids.data_ptr())
        torch.cuda.synchronize(); dt = time.time() - t
        if rep == 1: torch.cuda.profiler.stop()
    out[mode] = ids.clone()
    import ctypes
    fb = pkg.shn.lib().shn_debug_bruteforce_fallbacks; fb.restype = ctypes.c_ulonglong
    print(f"{mode}: fallbacks {fb() if mode == 'tc' else 0}", end="  ")
    print(f"{mode}: {dt:.3f}s  {2.0 * n * nq * dim / dt / 1e12:.1f} TFLOP/s (2*n*nq*d)", flush=True)
print("identical:", bool((out["simt"] == out["tc"]).all()))
