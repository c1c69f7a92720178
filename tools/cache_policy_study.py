"""Cache-policy study (CPU only, no GPU code involved): hit rates of the reference's compute-node cache policy
against the static, warm-up-chosen sets this repo replicates in HBM, on the SAME node-read traces under the reference's
Zipf query skew (scripts/data/skew.py recipe, tests/datagen.zipf_indices).

    python tools/cache_policy_study.py [n] [dim] [queries] > profiles/r2_cache_policy_study.json

Trace: HNSW::knn restated in numpy on a reference-format index (oracle build), logging every node read that goes through
HNSW::cache_lookup — the entry point (src/hnsw/hnsw.hh:263), upper-level candidates (:368, always admitted) and level-0
neighbours (:449, admitted with probability ADMISSION_RATIO = 0.01 once the cache is full).

The restated policy is pinned against the real one: tests/test_cache_policy.py runs the reference's own search path with --cache
(oracle/_ref) on the same index and stream — identical number of cache lookups, hit rate equal to within the eviction sampling
(e.g. 8 k nodes, 4 k Zipf(0) queries, 5 %: live 0.2353, simulated 0.2354 over 3 236 474 lookups).

Policies, each given the same number of entries E = cache_size / Node::size_until_components() as src/compute_node.cc:43-56
computes it from --cache-ratio:
  reference : src/cache/cache.hh:102-311 + cooling_table.hh:52-99 — hashed buckets, admission as above, eviction = pick a random
              entry, move it to the cooling table (FIFO buckets of 6, 10 % of the cache), evict what the cooling bucket drops unless it
              was touched again in the meantime (a hit on a cooling entry takes it out of the table)
  static    : this repo — all upper-level nodes plus the most-read level-0 nodes of a warm-up trace (the first 20 % of the queries,
              compute_node.cc:116-131 warm-up pass), never changed afterwards (shn_index_partition hot set)
  static/cn : with routing (k-means over the upper-level nodes, k = compute nodes; a query runs on the node of its nearest
              centroid, src/router/query_router.hh:356-368): every compute node keeps ITS OWN static set, chosen from the reads of the
              queries routed to it (hot set + shn_index_partition_build_halo), against the reference policy run per compute node
"""
import heapq
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import datagen  # noqa: E402
import hnsw_oracle  # noqa: E402

ADMISSION_RATIO = 0.01            # src/common/constants.hh:16
COOLING_BUCKET_ENTRIES = 6        # :14
COOLING_TABLE_RATIO = 0.1         # :15


def load_graph(base, m, efc):
    dump, _ = hnsw_oracle.build(base, m=m, efc=efc, seed=1234)
    ix = hnsw_oracle.Index([dump], base.shape[1], m)
    ex = ix.export()
    level, vec = ex["level"].astype(np.int64), ex["vectors"]
    l0 = ex["l0_adj"].astype(np.int64)
    l0[np.arange(l0.shape[1])[None, :] >= ex["l0_cnt"][:, None]] = -1
    upper = {}
    for r in np.flatnonzero(level > 0):
        for lv in range(1, int(level[r]) + 1):
            upper[(int(r), lv)] = np.asarray(ix.neighbors(int(r), lv), np.int64)
    return ix, vec, level, l0, upper


def knn_trace(q, vec, level, l0, upper, ep, ef):
    """HNSW::knn (hnsw.hh:253-307) with the reads that pass through cache_lookup logged in order: (node, is_inner)."""
    reads = [(ep, True)]
    d = lambda r: float(((vec[r] - q) ** 2).sum())
    cur, cd = ep, d(ep)
    for lv in range(int(level[ep]), 0, -1):          # search_for_one (:332-393)
        changed = True
        while changed:
            changed = False
            for nb in upper[(cur, lv)]:
                reads.append((int(nb), True))
                dn = d(nb)
                if dn < cd:
                    cur, cd, changed = int(nb), dn, True
    visited = {cur}
    cand = [(cd, cur)]                                # min-heap
    top = [(-cd, cur)]                                # max-heap, <= ef
    while cand:                                       # search_level (:407-476)
        dc, c = heapq.heappop(cand)
        if dc > -top[0][0]:
            break
        for nb in l0[c]:
            if nb < 0:
                break
            nb = int(nb)
            if nb in visited:
                continue
            visited.add(nb)
            reads.append((nb, False))
            dn = d(nb)
            if len(top) < ef or dn < -top[0][0]:
                heapq.heappush(cand, (dn, nb))
                heapq.heappush(top, (-dn, nb))
                if len(top) > ef:
                    heapq.heappop(top)
    return reads


class ReferenceCache:
    """cache.hh + cooling_table.hh restated for one compute node (single-threaded: the optimistic-lock retries vanish)."""

    def __init__(self, entries, seed):
        self.cap = max(1, entries)
        self.rng = np.random.default_rng(seed)
        self.keys = []                 # dense array of cached keys (random eviction candidate = random entry)
        self.pos = {}                  # key -> index in keys
        self.cooling = set()
        self.nb = max(1, int(np.ceil(self.cap / COOLING_BUCKET_ENTRIES * COOLING_TABLE_RATIO)))
        self.table = [[] for _ in range(self.nb)]   # FIFO buckets, newest first
        self.hits = self.misses = 0

    def _bucket(self, key):
        return (key * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF) % self.nb

    def _remove(self, key):
        i = self.pos.pop(key)
        last = self.keys.pop()
        if last != key:
            self.keys[i] = last
            self.pos[last] = i
        self.cooling.discard(key)

    def _evict_one(self):
        while True:                    # cache.hh:232-311
            key = self.keys[int(self.rng.integers(len(self.keys)))]
            if key in self.cooling:
                continue
            b = self.table[self._bucket(key)]
            victim = b.pop() if len(b) >= COOLING_BUCKET_ENTRIES else None   # cooling_table.hh:80-97
            b.insert(0, key)
            self.cooling.add(key)
            if victim is not None and victim in self.cooling:
                self._remove(victim)
                return

    def access(self, key, inner):
        if key in self.pos:            # cache.hh:102-145
            self.hits += 1
            if key in self.cooling:
                b = self.table[self._bucket(key)]
                if key in b:
                    b.remove(key)
                    self.cooling.discard(key)
            return
        self.misses += 1
        full = len(self.keys) >= self.cap
        admit = True if inner else (True if not full else self.rng.random() < ADMISSION_RATIO)   # hnsw.hh:368,448
        if not admit:
            return
        if full:
            self._evict_one()
        self.pos[key] = len(self.keys)
        self.keys.append(key)


def static_set(traces, level, entries):
    counts = {}
    for tr in traces:
        for node, _ in tr:
            counts[node] = counts.get(node, 0) + 1
    inner = [int(r) for r in np.flatnonzero(level > 0)]
    chosen = set(inner[:entries])
    rest = sorted((r for r in counts if r not in chosen), key=lambda r: (-counts[r], r))
    for r in rest:
        if len(chosen) >= entries:
            break
        chosen.add(r)
    return chosen


def kmeans(x, k, seed, iters=25):
    rng = np.random.default_rng(seed)
    c = x[rng.choice(len(x), k, replace=False)].copy()
    for _ in range(iters):
        a = ((x[:, None, :] - c[None, :, :]) ** 2).sum(2).argmin(1)
        for j in range(k):
            if (a == j).any():
                c[j] = x[a == j].mean(0)
    return c


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    m, efc, ef, pool_size, cns = 16, 100, 64, 5000, 4
    base = datagen.latent_rows(n, dim, 1001)
    pool = datagen.latent_rows(pool_size, dim, 2002)
    ix, vec, level, l0, upper = load_graph(base, m, efc)
    ep = ix.entry_row
    node_bytes = 16 + 4 * dim
    index_bytes = sum(hnsw_oracle_alloc(dim, m, int(l)) for l in level)
    pool_traces = [knn_trace(pool[i], vec, level, l0, upper, ep, ef) for i in range(pool_size)]
    cent = kmeans(vec[level > 0], cns, 7)
    pool_dest = ((pool[:, None, :] - cent[None, :, :]) ** 2).sum(2).argmin(1)
    out = dict(n=n, dim=dim, m=m, ef=ef, queries=nq, distinct_queries=pool_size, compute_nodes_routed=cns,
               reads_per_query=round(float(np.mean([len(t) for t in pool_traces])), 1),
               inner_nodes_pct=round(100.0 * float((level > 0).mean()), 2), rows=[])
    for alpha in (0.0, 0.5, 1.0):
        order = datagen.zipf_indices(pool_size, nq, alpha, seed=99)
        warm = order[: nq // 5]
        run = order[nq // 5:]
        for ratio in (2, 5, 10):
            entries = int(index_bytes / 100.0 * ratio / node_bytes)      # compute_node.cc:43-54
            # one compute node sees every query
            ref = ReferenceCache(entries, 1)
            for qi in warm:                                              # the warm-up pass fills the cache (compute_node.cc:116-131)
                for node, inner in pool_traces[qi]:
                    ref.access(node, inner)
            ref.hits = ref.misses = 0
            for qi in run:
                for node, inner in pool_traces[qi]:
                    ref.access(node, inner)
            hot = static_set([pool_traces[qi] for qi in warm], level, entries)
            sh = sm = 0
            for qi in run:
                for node, _ in pool_traces[qi]:
                    if node in hot:
                        sh += 1
                    else:
                        sm += 1
            # routed: every compute node runs the policy over the queries routed to it
            rh = rm = th = tm = 0
            for cn in range(cns):
                w = [qi for qi in warm if pool_dest[qi] == cn]
                r = [qi for qi in run if pool_dest[qi] == cn]
                rc = ReferenceCache(entries, 10 + cn)
                for qi in w:
                    for node, inner in pool_traces[qi]:
                        rc.access(node, inner)
                rc.hits = rc.misses = 0
                for qi in r:
                    for node, inner in pool_traces[qi]:
                        rc.access(node, inner)
                rh += rc.hits; rm += rc.misses
                own = static_set([pool_traces[qi] for qi in w], level, entries)
                for qi in r:
                    for node, _ in pool_traces[qi]:
                        if node in own:
                            th += 1
                        else:
                            tm += 1
            out["rows"].append(dict(zipf_alpha=alpha, cache_ratio_pct=ratio, entries=entries,
                                    one_cn=dict(reference=round(ref.hits / (ref.hits + ref.misses), 4), static=round(sh / (sh + sm), 4)),
                                    routed_4cn=dict(reference=round(rh / (rh + rm), 4), static_per_cn=round(th / (th + tm), 4))))
            print(out["rows"][-1], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


def hnsw_oracle_alloc(dim, m, level):
    s = (16 + 4 * dim) + (4 + 8 * 2 * m) + level * (4 + 8 * m)   # node.hh:45-53, rdma_atomics.hh:88-95
    while s % 8:
        s += 4
    return s


if __name__ == "__main__":
    main()
