"""The synthetic-data recipes: the Zipf query skew of the reference's scripts/data/skew.py."""
import numpy as np

import datagen


def test_zipf_matches_the_reference_recipe():
    n, pool, alpha = 5000, 1000, 1.0
    idx = datagen.zipf_indices(pool, n, alpha)
    assert len(idx) == n and idx.min() == 0
    counts = np.bincount(idx, minlength=pool)
    h = (np.arange(1, pool + 1) ** -alpha).sum()
    want = np.ceil(n * (np.arange(1, pool + 1) ** -alpha) / h)
    used = np.nonzero(counts)[0].max()
    assert (counts[:used] == want[:used]).all() and counts[used] <= want[used] and (counts[used + 1:] == 0).all()
    assert (np.diff(counts[:used + 1]) <= 0).all()  # most popular first, as in pool order
    assert not (np.diff(idx) >= 0).all()            # shuffled
    assert (datagen.zipf_indices(pool, n, alpha) == idx).all()  # seeded


def test_zipf_alpha_zero_is_uniform_prefix():
    idx = datagen.zipf_indices(1000, 500, 0.0)
    assert sorted(idx.tolist()) == list(range(500))  # ceil(n * 1/pool) = 1 occurrence each until n are drawn


def test_latent_rows_are_reproducible():
    a = datagen.latent_rows(10, 24, 5)
    b = datagen.latent_rows(10, 24, 5)
    assert (a == b).all() and a.dtype == np.float32
    c = datagen.latent_rows(10, 24, 5, normalize=True)
    assert np.allclose(np.linalg.norm(c, axis=1), 1, atol=1e-6)
