"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (oracle/_ref/libshine_ref.so: /root/reference's
src/hnsw/hnsw.hh insert/knn compiled unmodified, see oracle/ref_harness.cc).  Run in the build container:
    python tests/golden/make_golden.py
Each fixture holds the base rows, the queries, the index dump(s) the reference's memory node(s) would write, and for
several (k, ef) the ids the reference returns (heap-array order, hnsw.hh:300-303), the distance of each id under the
reference's own Distance::dist, and the reference's per-query counters."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import datagen  # noqa: E402
import shine_ref  # noqa: E402

CASES = [
    # name, n, nq, dim, m, efc, ip, num_mn, normalize, dup, runs
    ("l2_d32_n2000_m16", 2000, 100, 32, 16, 100, False, 1, False, 0, [(10, 64), (10, 16), (1, 1), (10, 10), (5, 200)]),
    ("l2_d128_n1500_m16", 1500, 64, 128, 16, 100, False, 1, False, 0, [(10, 64), (10, 32), (20, 128)]),
    ("ip_d40_n1500_m8", 1500, 64, 40, 8, 80, True, 1, True, 0, [(10, 64), (10, 20)]),
    ("l2_d96_n1500_m16_2mn", 1500, 64, 96, 16, 100, False, 2, True, 0, [(10, 64), (10, 100)]),
    ("ip_d200_n1000_m16", 1000, 48, 200, 16, 100, True, 1, True, 0, [(10, 64), (10, 250)]),
    ("l2_d20_n300_m4_3mn", 300, 40, 20, 4, 40, False, 3, False, 0, [(10, 64), (3, 8), (10, 400)]),
    ("l2_d16_n600_m8_dups", 600, 60, 16, 8, 60, False, 1, False, 150, [(10, 64), (10, 12)]),
    ("l2_d960_n400_m16", 400, 16, 960, 16, 60, False, 1, False, 0, [(10, 32)]),
    ("l2_d8_n40_m32", 40, 10, 8, 32, 40, False, 1, False, 0, [(10, 64), (10, 10)]),
]


def main():
    for name, n, nq, dim, m, efc, ip, num_mn, normalize, dup, runs in CASES:
        base, queries = datagen.base_and_queries(n, nq, dim, normalize=normalize)
        if dup:  # exact duplicates -> exact distance ties
            base[n - dup:] = base[:dup]
            queries[: nq // 4] = base[: nq // 4]
        dumps, bstats, _ = shine_ref.build(base, m=m, efc=efc, seed=1234, ip=ip, threads=1, coroutines=4, num_mn=num_mn)
        out = dict(dim=dim, m=m, efc=efc, ip=ip, n=n, num_mn=num_mn, base=base, queries=queries)
        for i, d in enumerate(dumps):
            out[f"dump{i}"] = np.frombuffer(d, dtype=np.uint8)
        for k, ef in runs:
            ids, dists, counts, st, _ = shine_ref.search(dumps, dim, m, queries, k, ef, ip=ip, per_query_stats=True)
            tag = f"{k}_{ef}"
            out[f"ids_{tag}"], out[f"dists_{tag}"], out[f"counts_{tag}"] = ids, dists, counts
            for s in ("distcomps", "visited_nodes", "visited_nodes_l0", "visited_neighborlists", "rdma_reads_in_bytes"):
                out[f"stat_{tag}_{s}"] = st[s]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        gt = datagen.bruteforce(base, queries, 10, ip=ip)
        k0, ef0 = runs[0]
        ids0 = out[f"ids_{k0}_{ef0}"]
        print(f"{name}: dump bytes {[len(d) for d in dumps]}, build distcomps {bstats['distcomps']}, "
              f"recall@{k0} ef={ef0}: {datagen.recall(ids0[:, :min(k0, 10)], gt[:, :min(k0, 10)]):.3f}, "
              f"file {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
