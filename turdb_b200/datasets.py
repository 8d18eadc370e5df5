"""Seeded synthetic corpora for the BASELINE.json configs.

Pure numpy; used by tests/ and bench.py so every result file can name (generator, seed).
"""
from __future__ import annotations

import numpy as np


def _normalise(x: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(x, axis=1, keepdims=True)
    n[n == 0] = 1.0
    return (x / n).astype(np.float32)


def gaussian_latent(n: int, dim: int, seed: int, latent: int = 16, noise: float = 0.05,
                    normalise: bool = False, basis_seed: int = 7, chunk: int = 1 << 18) -> np.ndarray:
    """x = A z + noise * eps, z ~ N(0, I_latent), A ~ N(0, 1/latent).  `basis_seed` fixes A so that
    corpus and queries (different `seed`) share the same manifold."""
    a = np.random.default_rng(basis_seed).normal(0.0, 1.0 / np.sqrt(latent), (latent, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = rng.standard_normal((e - s, latent), dtype=np.float32)
        x = z @ a + noise * rng.standard_normal((e - s, dim), dtype=np.float32)
        out[s:e] = _normalise(x) if normalise else x
    return out


def iid_gaussian(n: int, dim: int, seed: int, normalise: bool = False) -> np.ndarray:
    x = np.random.default_rng(seed).standard_normal((n, dim), dtype=np.float32)
    return _normalise(x) if normalise else x


def clustered(n: int, dim: int, seed: int, n_centres: int | None = None, sigma: float = 0.1,
              centre_seed: int = 11, normalise: bool = False, chunk: int = 1 << 18, centre_latent: int = 0,
              corpus_n: int | None = None) -> np.ndarray:
    """1000 * (corpus_n / 100k) centres, isotropic sigma around a uniformly chosen centre.
    centre_latent == 0: centres ~ N(0, I_dim).  With tens of thousands of such centres the
    centres themselves are an i.i.d. high-dimensional point set, on which no graph index reaches recall 0.95
    (SURVEY's probe: 100k x 128 i.i.d. Gaussian tops out at 0.87 with ef 512) — reported, labelled.
    centre_latent > 0: centres drawn from the latent-Gaussian manifold (x = A z, z ~ N(0, I_latent)), i.e.
    clusters along a low-dimensional structure as in real embedding corpora.
    `corpus_n` sizes the centre set (pass the corpus size when generating queries)."""
    if n_centres is None:
        n_centres = max(10, int(1000 * (corpus_n or n) / 100_000))
    crng = np.random.default_rng(centre_seed)
    if centre_latent:
        a = crng.normal(0.0, 1.0 / np.sqrt(centre_latent), (centre_latent, dim)).astype(np.float32)
        centres = (crng.standard_normal((n_centres, centre_latent), dtype=np.float32) @ a).astype(np.float32)
    else:
        centres = crng.standard_normal((n_centres, dim), dtype=np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        c = rng.integers(0, n_centres, e - s)
        x = centres[c] + sigma * rng.standard_normal((e - s, dim), dtype=np.float32)
        out[s:e] = _normalise(x) if normalise else x
    return out


def sift_like(n: int, dim: int, seed: int, latent: int = 16, basis_seed: int = 13, chunk: int = 1 << 18) -> np.ndarray:
    """SIFT-shaped: non-negative integer-valued floats in 0..218 with SIFT-like low intrinsic dimension
    (a latent-Gaussian manifold quantised to integers, so squared distances are exact integers with ties)."""
    a = np.random.default_rng(basis_seed).normal(0.0, 1.0 / np.sqrt(latent), (latent, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        z = rng.standard_normal((e - s, latent), dtype=np.float32)
        y = z @ a + 0.05 * rng.standard_normal((e - s, dim), dtype=np.float32)
        out[s:e] = np.clip(np.rint(60.0 + 40.0 * y), 0, 218)
    return out


GENERATORS = {
    "gaussian_latent": gaussian_latent,
    "iid_gaussian": iid_gaussian,
    "clustered": clustered,
    "sift_like": sift_like,
}


def make(generator: str, n: int, dim: int, seed: int, **kw) -> np.ndarray:
    return GENERATORS[generator](n, dim, seed, **kw)
