"""Pins the CPU restatement (oracle/hnsw_oracle.c) to the reference: against the committed golden vectors, which
were produced by the reference's own code (tests/golden/make_golden.py), and — where oracle/_ref/libshine_ref.so is
present — against the reference run live on fresh inputs.  Bit-exact: ids in the reference's heap-array order,
distances, and the reference's counters."""
import numpy as np
import pytest

import datagen
import golden_io
import hnsw_oracle
import shine_ref

CASES = golden_io.case_names()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    case = golden_io.load_case(name)
    ix = hnsw_oracle.Index(case["dumps"], case["dim"], case["m"])
    assert ix.n == case["n"]
    for (k, ef), run in case["runs"].items():
        ids, dists, counts, ct = ix.knn(case["queries"], k, ef, ip=case["ip"], counters=True)
        assert (counts == run["counts"]).all()
        assert (ids == run["ids"]).all(), f"{name} k={k} ef={ef}: ids (heap-array order) differ from the reference"
        assert (dists.view(np.uint32) == run["dists"].view(np.uint32)).all(), "distance bits differ"
        st = run["stats"]
        assert (ct["distcomps"] == st["distcomps"]).all()
        assert (ct["visited_nodes"] == st["visited_nodes"]).all()
        assert (ct["visited_nodes_l0"] == st["visited_nodes_l0"]).all()
        assert (ct["lists_l0"] + ct["lists_upper"] == st["visited_neighborlists"]).all()
        # the reference also READs the 8-byte entry-point pointer once per coroutine (rdma_reads.hh:74-99);
        # the fixture ran one fresh coroutine per query
        assert (ct["rdma_reads_in_bytes"] + 8 == st["rdma_reads_in_bytes"]).all()


@pytest.mark.parametrize("name", CASES)
def test_list_split_formula(name):
    """lists_upper can be solved from the reference's own counters; the oracle counts it directly."""
    case = golden_io.load_case(name)
    dim, m = case["dim"], case["m"]
    ix = hnsw_oracle.Index(case["dumps"], dim, m)
    (k, ef), run = next(iter(case["runs"].items()))
    _, _, _, ct = ix.knn(case["queries"], k, ef, ip=case["ip"], counters=True)
    st = run["stats"]
    lists = st["visited_neighborlists"].astype(np.int64)
    reads = st["rdma_reads_in_bytes"].astype(np.int64) - 8
    solved = ((4 + 16 * m) * lists + (16 + 4 * dim) * (st["distcomps"].astype(np.int64) - 1) - reads) // (8 * m)
    assert (solved == ct["lists_upper"].astype(np.int64)).all()


def test_multithreaded_oracle_is_deterministic():
    case = golden_io.load_case("l2_d32_n2000_m16")
    ix = hnsw_oracle.Index(case["dumps"], case["dim"], case["m"])
    a = ix.knn(case["queries"], 10, 64, threads=1)
    b = ix.knn(case["queries"], 10, 64, threads=4)
    assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()


def test_select_heuristic_small():
    rng = np.random.default_rng(5)
    vec = rng.standard_normal((40, 24)).astype(np.float32)
    q = rng.standard_normal(24).astype(np.float32)
    d = np.array([hnsw_oracle.dist(q, v) for v in vec], np.float32)
    sel, dc = hnsw_oracle.select_heuristic(np.arange(40), d, vec, 8)
    assert 1 <= len(sel) <= 8 and sel[0] == int(np.argmin(d))
    # every kept candidate is closer to the query than to any earlier kept one (hnsw.hh:495-518)
    for i, c in enumerate(sel):
        for s in sel[:i]:
            assert hnsw_oracle.dist(vec[s], vec[c]) >= d[c]
    few, _ = hnsw_oracle.select_heuristic(np.arange(5), d[:5], vec[:5], 8)
    assert list(few) == [0, 1, 2, 3, 4]  # fewer than m candidates: all kept (hnsw.hh:483)


needs_ref = pytest.mark.skipif(not shine_ref.available(), reason="oracle/_ref/libshine_ref.so not built")


@needs_ref
@pytest.mark.parametrize("dim,ip", [(24, False), (128, False), (50, True), (7, False)])
def test_distance_bits_match_reference_live(dim, ip):
    rng = np.random.default_rng(dim)
    for _ in range(200):
        a = rng.standard_normal(dim).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32)
        assert np.float32(hnsw_oracle.dist(a, b, ip)).view(np.uint32) == np.float32(shine_ref.dist(a, b, ip)).view(np.uint32)


@needs_ref
@pytest.mark.parametrize("n,dim,m,ip,num_mn", [(1200, 48, 12, False, 1), (900, 36, 6, True, 2)])
def test_oracle_matches_reference_live(n, dim, m, ip, num_mn):
    """Fresh inputs (not in the fixtures): build with the reference, search with both."""
    base, queries = datagen.base_and_queries(n, 50, dim, normalize=ip)
    base = base[::-1].copy()  # not the fixture rows
    dumps, _, _ = shine_ref.build(base, m=m, efc=60, seed=7, ip=ip, num_mn=num_mn)
    ix = hnsw_oracle.Index(dumps, dim, m)
    for k, ef in [(10, 40), (1, 5)]:
        rid, rd, rc, st, _ = shine_ref.search(dumps, dim, m, queries, k, ef, ip=ip, per_query_stats=True)
        oid, od, oc, ct = ix.knn(queries, k, ef, ip=ip, counters=True)
        assert (rid == oid).all() and (rc == oc).all()
        assert (rd.view(np.uint32) == od.view(np.uint32)).all()
        assert (st["distcomps"] == ct["distcomps"]).all()


@needs_ref
def test_reference_cache_does_not_change_results():
    """hnsw.hh:525-548: the compute-node cache only changes bytes moved, never a result."""
    case = golden_io.load_case("l2_d32_n2000_m16")
    a = shine_ref.search(case["dumps"], 32, 16, case["queries"], 10, 64, cache_ratio_pct=0)
    b = shine_ref.search(case["dumps"], 32, 16, case["queries"], 10, 64, cache_ratio_pct=20)
    assert (a[0] == b[0]).all()
    assert b[3]["cache_hits"] > 0


@needs_ref
@pytest.mark.parametrize("c,dim,m,ip", [(40, 24, 8, False), (5, 16, 8, False), (200, 128, 16, False), (64, 40, 32, True),
                                        (16, 32, 16, False), (33, 96, 32, False)])
def test_select_heuristic_matches_reference_live(c, dim, m, ip):
    """The reference's own HNSW<D>::select_heuristic (hnsw.hh:482-522, reached through the harness) vs the restatement:
    same selected candidates, same distance-computation count."""
    rng = np.random.default_rng(c * 1000 + dim)
    vec = rng.standard_normal((c, dim)).astype(np.float32)
    q = rng.standard_normal(dim).astype(np.float32)
    d = np.array([shine_ref.dist(q, v, ip) for v in vec], np.float32)
    uids = rng.permutation(10 * c)[:c].astype(np.uint32)
    r_sel, r_dc = shine_ref.select_heuristic(uids, d, vec, m, ip)
    o_sel, o_dc = hnsw_oracle.select_heuristic(uids, d, vec, m, ip)
    assert sorted(r_sel.tolist()) == sorted(o_sel.tolist()) and r_dc == o_dc


@needs_ref
@pytest.mark.parametrize("n,dim,m,efc,ip", [(300, 16, 4, 20, False), (2000, 32, 16, 100, False), (1500, 40, 8, 60, True),
                                            (700, 128, 32, 50, False), (1, 8, 4, 10, False), (12, 8, 2, 10, False)])
def test_build_restatement_is_the_reference_build(n, dim, m, efc, ip):
    """orc_build (HNSW::insert restated) against the reference's own single-thread / single-coroutine build, which is
    deterministic: identical levels, identical level-0 and upper lists IN STORED ORDER, identical entry point, identical
    dump size and distance-computation count."""
    base, queries = datagen.base_and_queries(n, 20, dim, normalize=ip)
    ref, rst, _ = shine_ref.build(base, m=m, efc=efc, seed=1234, ip=ip, threads=1, coroutines=1)
    mine, dc = hnsw_oracle.build(base, m=m, efc=efc, seed=1234, ip=ip)
    assert len(mine) == len(ref[0]) and dc == rst["distcomps"]
    a, b = hnsw_oracle.Index(ref, dim, m), hnsw_oracle.Index([mine], dim, m)
    assert a.entry_row == b.entry_row and a.max_level == b.max_level
    ea, eb = a.export(), b.export()
    for key in ea:
        assert (ea[key] == eb[key]).all(), key
    for r in np.nonzero(ea["level"] > 0)[0]:
        for l in range(1, ea["level"][r] + 1):
            assert np.array_equal(a.neighbors(int(r), l), b.neighbors(int(r), l))
    if n >= 100:  # and therefore the same search results
        ra = a.knn(queries, 10, 50, ip=ip)
        rb = b.knn(queries, 10, 50, ip=ip)
        assert (ra[0] == rb[0]).all()
