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
