"""tools/cache_policy_study.py's restatement of the reference's compute-node cache (src/cache/cache.hh + cooling_table.hh,
admission src/hnsw/hnsw.hh:368,448) is pinned here against the REAL one: the reference's own search path, compiled unmodified
(oracle/_ref), run with --cache on the same index and the same query stream.  The number of reads that pass through
HNSW::cache_lookup must be identical (the numpy HNSW::knn of the study logs exactly the reads the reference performs) and the hit
rate must agree up to the randomness of the eviction sampling.  CPU only."""
import os
import sys

import numpy as np
import pytest

import datagen
import hnsw_oracle
import shine_ref

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import cache_policy_study as cps  # noqa: E402


@pytest.mark.skipif(not shine_ref.available(), reason="oracle/_ref not built")
def test_simulated_reference_cache_matches_the_reference_live():
    n, dim, m, efc, ef = 3000, 24, 16, 60, 48
    base = datagen.latent_rows(n, dim, 1001)
    pool = datagen.latent_rows(150, dim, 2002)
    dump, _ = hnsw_oracle.build(base, m=m, efc=efc, seed=1234)
    ix = hnsw_oracle.Index([dump], dim, m)
    ex = ix.export()
    level, vec = ex["level"].astype(np.int64), ex["vectors"]
    l0 = ex["l0_adj"].astype(np.int64)
    l0[np.arange(l0.shape[1])[None, :] >= ex["l0_cnt"][:, None]] = -1
    upper = {(int(r), lv): np.asarray(ix.neighbors(int(r), lv), np.int64)
             for r in np.flatnonzero(level > 0) for lv in range(1, int(level[r]) + 1)}
    traces = [cps.knn_trace(q, vec, level, l0, upper, ix.entry_row, ef) for q in pool]
    index_bytes = sum(cps.hnsw_oracle_alloc(dim, m, int(l)) for l in level)
    order = datagen.zipf_indices(len(pool), 600, 1.0, seed=99)
    for ratio in (10, 25):
        _, _, _, st, _ = shine_ref.search([dump], dim, m, pool[order], 10, ef, threads=1, coroutines=1, cache_ratio_pct=ratio)
        live_total = int(st["cache_hits"] + st["cache_misses"])
        live = st["cache_hits"] / live_total
        sim = cps.ReferenceCache(int(index_bytes / 100.0 * ratio / (16 + 4 * dim)), 3)
        for qi in order:
            for node, inner in traces[qi]:
                sim.access(node, inner)
        assert sim.hits + sim.misses == live_total, "the study's trace is not the reference's sequence of cache lookups"
        assert abs(sim.hits / live_total - live) < 0.02, (ratio, live, sim.hits / live_total)
