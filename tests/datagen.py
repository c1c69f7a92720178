"""Synthetic datasets of the shapes BASELINE.json names: a low-intrinsic-dimension latent model so
that HNSW behaves as on real descriptors.  x = A z / sqrt(r) + 0.05 eps, A ~ N(0,1)^{d x r} (seed 1), r = 16."""
import numpy as np


def latent_rows(n, dim, seed, r=16, normalize=False, noise=0.05):
    a = np.random.default_rng(1).standard_normal((dim, r)).astype(np.float32)
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, r)).astype(np.float32)
    eps = rng.standard_normal((n, dim)).astype(np.float32)
    x = z @ a.T / np.float32(np.sqrt(r)) + np.float32(noise) * eps
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32)


def base_and_queries(n, nq, dim, normalize=False):
    return latent_rows(n, dim, 1001, normalize=normalize), latent_rows(nq, dim, 2002, normalize=normalize)


def bruteforce(base, queries, k, ip=False):
    """Exact top-k in float64 (ids ascending by distance, ties by id)."""
    b = base.astype(np.float64)
    q = queries.astype(np.float64)
    if ip:
        d = 1.0 - q @ b.T
    else:
        d = (q * q).sum(1)[:, None] - 2.0 * (q @ b.T) + (b * b).sum(1)[None, :]
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    return idx.astype(np.uint32)


def recall(ids, gt):
    k = gt.shape[1]
    hit = 0
    for a, b in zip(ids, gt):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / (gt.shape[0] * k)


def zipf_indices(pool_size, n, alpha, seed=1234):
    """Query skew exactly as scripts/data/skew.py:114-153 of the reference: pool entry k (1-based, pool order) is drawn
    ceil(n * p_k) times with p_k = k^-alpha / H(pool_size, alpha) until n are drawn, then the sequence is shuffled — with a
    SEEDED permutation (the reference's is unseeded, :153).  alpha = 0 is the uniform distribution."""
    k = np.arange(1, pool_size + 1, dtype=np.float64)
    p = k ** -float(alpha)
    p /= p.sum()
    occ = np.ceil(n * p).astype(np.int64)
    cum = np.cumsum(occ)
    last = int(np.searchsorted(cum, n))            # first pool entry at which >= n have been drawn
    occ = occ[: last + 1].copy()
    occ[-1] -= int(cum[last] - n)                    # the reference asserts drawn == n; trim the overshoot instead
    idx = np.repeat(np.arange(last + 1, dtype=np.int64), occ)
    assert len(idx) == n
    return idx[np.random.default_rng(seed).permutation(n)]
