"""shn_bruteforce_topk: exact ground truth on the GPU vs a float64 numpy reference."""
import numpy as np
import pytest

import datagen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nq,dim,k,ip", [(5000, 70, 32, 10, False), (3001, 1, 128, 100, False), (2000, 65, 40, 7, True),
                                           (129, 200, 8, 20, False), (50, 3, 200, 64, True), (100000, 300, 96, 10, False)])
def test_matches_float64(pkg, n, nq, dim, k, ip):
    base, queries = datagen.base_and_queries(n, nq, dim, normalize=ip)
    ids, dists = pkg.bruteforce_topk(base, queries, k, ip=ip)
    kk = min(k, n)
    b, q = base.astype(np.float64), queries.astype(np.float64)
    exact = 1.0 - q @ b.T if ip else ((q[:, None, :] - b[None, :, :]) ** 2).sum(2) if n * nq < 2e6 else \
        (q * q).sum(1)[:, None] - 2 * q @ b.T + (b * b).sum(1)[None, :]
    want = np.sort(exact, axis=1)[:, :kk]
    assert (ids[:, :kk] < n).all() and (ids[:, kk:] == 0xFFFFFFFF).all()
    got_exact = np.take_along_axis(exact, ids[:, :kk].astype(np.int64), 1)
    assert np.allclose(got_exact, want, rtol=1e-5, atol=1e-5), "not the k nearest"
    assert np.allclose(dists[:, :kk], want, rtol=1e-4, atol=1e-5)
    assert (np.diff(dists[:, :kk], axis=1) >= 0).all()
    for row in ids[:, :kk]:
        assert len(set(row.tolist())) == kk


def test_ties_break_by_lower_id(pkg):
    base = np.zeros((300, 16), np.float32)
    base[100:] = 1.0
    ids, dists = pkg.bruteforce_topk(base, np.zeros((2, 16), np.float32), 10)
    assert (ids == np.arange(10)).all() and (dists == 0).all()


def test_recall_of_the_index_against_gpu_ground_truth(pkg):
    n, dim = 20000, 64
    base, queries = datagen.base_and_queries(n, 500, dim)
    gt, _ = pkg.bruteforce_topk(base, queries, 10)
    assert (gt == datagen.bruteforce(base, queries, 10)).mean() > 0.999
    with pkg.Index.build(base, 16, 200) as ix:
        ids, _, _ = ix.search(queries, 10, 64)
    assert datagen.recall(ids, gt) > 0.98


@pytest.mark.parametrize("n,nq,dim,k,ip", [(1000, 5, 64, 10, False), (70000, 300, 128, 10, False), (33333, 129, 192, 16, True),
                                           (5000, 64, 960, 10, False), (300000, 1000, 128, 10, False),
                                           (50000, 200, 96, 10, False), (40000, 150, 200, 10, True), (3000, 40, 40, 5, False)])
def test_tensor_core_path_is_exact(pkg, monkeypatch, n, nq, dim, k, ip):
    """tcgen05 candidate generation (split-bf16, 3 products, K zero-padded to a multiple of 64) + fp32 re-rank returns what the
    fp32-pipe kernel returns: same ids, same distance bits (both rank with the same arithmetic; the tensor cores only pick
    candidates, and a query whose candidates cannot be certified complete is answered by the fp32 kernel)."""
    import ctypes
    base, queries = datagen.base_and_queries(n, nq, dim, normalize=ip)
    monkeypatch.setenv("SHN_BRUTEFORCE", "simt")
    ids_s, d_s = pkg.bruteforce_topk(base, queries, k, ip=ip)
    monkeypatch.setenv("SHN_BRUTEFORCE", "tc")
    ids_t, d_t = pkg.bruteforce_topk(base, queries, k, ip=ip)
    same = (ids_s == ids_t).all(axis=1)
    assert same.mean() == 1.0, f"{(~same).sum()} of {nq} queries differ"
    assert (d_s.view(np.uint32) == d_t.view(np.uint32)).all()
    fb = pkg.shn.lib().shn_debug_bruteforce_fallbacks
    fb.restype = ctypes.c_ulonglong
    assert fb() <= nq // 10, "the certificate should hold for nearly every query of a well-spread dataset"


def test_tensor_core_certificate_catches_what_the_candidates_miss(pkg, monkeypatch):
    """80 near-copies of every query (distances far below the split-bf16 error of the candidate scores) in ONE slice: the 32
    candidates of a slice that holds more than 32 of them cannot be proven to contain the 10 nearest, the certificate fails, and the fp32 kernel answers —
    the result is still exactly the fp32-pipe kernel's."""
    import ctypes
    rng = np.random.default_rng(5)
    nq, dim, k = 8, 128, 10
    queries = (rng.standard_normal((nq, dim)) * 30).astype(np.float32)
    near = np.repeat(queries, 80, axis=0) + rng.standard_normal((nq * 80, dim)).astype(np.float32) * 1e-4
    far = (rng.standard_normal((2000, dim)) * 30).astype(np.float32)
    base = np.ascontiguousarray(np.concatenate([near, far]), dtype=np.float32)
    monkeypatch.setenv("SHN_BRUTEFORCE", "simt")
    ids_s, d_s = pkg.bruteforce_topk(base, queries, k)
    monkeypatch.setenv("SHN_BRUTEFORCE", "tc")
    ids_t, d_t = pkg.bruteforce_topk(base, queries, k)
    assert (ids_s == ids_t).all() and (d_s.view(np.uint32) == d_t.view(np.uint32)).all()
    fb = pkg.shn.lib().shn_debug_bruteforce_fallbacks
    fb.restype = ctypes.c_ulonglong
    assert fb() == nq


def test_tensor_core_path_rejects_unsupported_shapes(pkg, monkeypatch):
    monkeypatch.setenv("SHN_BRUTEFORCE", "tc")
    base, queries = datagen.base_and_queries(100, 4, 40)
    with pytest.raises(pkg.ShnError):
        pkg.bruteforce_topk(base, queries, 17)   # k above CAND / 2: no margin for the certificate
