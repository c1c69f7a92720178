"""Scratch: recall of GPU-built vs reference-built graphs for several batch policies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import __graft_entry__ as ge
import datagen, shine_ref
pkg = ge.load_package()
n, dim, m, efc = int(sys.argv[1]), int(sys.argv[2]), 16, 200
base, queries = datagen.base_and_queries(n, 2000, dim)
gt = datagen.bruteforce(base, queries, 10)
efs = (10, 16, 32, 64, 128)
ref_dumps, _, secs = shine_ref.build(base, m=m, efc=efc, threads=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
with pkg.Index.from_dumps(ref_dumps, dim, m) as ix:
    print(f"reference-built ({secs:.0f}s):", [round(datagen.recall(ix.search(queries, 10, ef)[0], gt), 4) for ef in efs], flush=True)
for div, bmax in ((8, 16384), (32, 16384), (64, 16384), (128, 16384), (32, 2048), (64, 512)):
    pkg.set_build_option("batch_div", div); pkg.set_build_option("batch_max", bmax)
    with pkg.Index.build(base, m, efc) as ix:
        bs = ix.build_stats()
        print(f"gpu div={div} bmax={bmax} ({bs['kernel_ms'] / 1e3:.2f}s):", [round(datagen.recall(ix.search(queries, 10, ef)[0], gt), 4) for ef in efs], flush=True)
