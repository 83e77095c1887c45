"""Seeded synthetic bitmap generators for parity tests (numpy only; no reference needed).

Each generator returns a sorted unique uint32 array.  The per-key "profiles" are chosen so
that, after roaring_bitmap_of_ptr (+/- run_optimize), every container type and every edge
of the reference's type rules is hit: tiny/mid/near-4096 arrays, >4096 bitsets, near-full and
full containers, few long runs, many short runs, single values at 0 / 65535.
"""
import numpy as np

PROFILES = ["tiny", "array", "edge4096", "bitset", "dense", "nearfull", "full",
            "longruns", "shortruns", "ends", "halves", "stripes"]


def container_values(rng, profile):
    if profile == "tiny":
        return rng.choice(65536, size=rng.integers(1, 40), replace=False)
    if profile == "array":
        return rng.choice(65536, size=rng.integers(40, 4000), replace=False)
    if profile == "edge4096":
        return rng.choice(65536, size=rng.integers(4090, 4103), replace=False)
    if profile == "bitset":
        return rng.choice(65536, size=rng.integers(4200, 30000), replace=False)
    if profile == "dense":
        return np.flatnonzero(rng.random(65536) < rng.uniform(0.4, 0.95))
    if profile == "nearfull":
        v = np.ones(65536, bool)
        v[rng.choice(65536, size=rng.integers(1, 6), replace=False)] = False
        return np.flatnonzero(v)
    if profile == "full":
        return np.arange(65536)
    if profile == "longruns":
        v = np.zeros(65536, bool)
        for _ in range(rng.integers(1, 12)):
            s = rng.integers(0, 65536)
            v[s:s + rng.integers(1, 9000)] = True
        return np.flatnonzero(v)
    if profile == "shortruns":
        v = np.zeros(65536, bool)
        n = rng.integers(20, 2500)
        starts = rng.choice(65536, size=n, replace=False)
        lens = rng.integers(1, 12, size=n)
        for s, l in zip(starts, lens):
            v[s:s + l] = True
        return np.flatnonzero(v)
    if profile == "ends":
        return np.array([0, 65535]) if rng.random() < 0.5 else np.array([65535])
    if profile == "halves":
        return np.arange(0, 65536, 2) if rng.random() < 0.5 else np.arange(1, 65536, 2)
    if profile == "stripes":
        step = int(rng.integers(2, 70))
        return np.arange(int(rng.integers(0, step)), 65536, step)
    raise ValueError(profile)


def random_bitmap(rng, n_keys=8, key_space=12, profiles=None):
    """Sorted unique uint32 values spread over `n_keys` of the first `key_space` keys."""
    profiles = profiles or PROFILES
    keys = np.sort(rng.choice(key_space, size=min(n_keys, key_space), replace=False))
    parts = []
    for k in keys:
        p = profiles[rng.integers(0, len(profiles))]
        v = np.sort(np.asarray(container_values(rng, p), dtype=np.uint32))
        parts.append((np.uint32(k) << np.uint32(16)) | v)
    if not parts:
        return np.zeros(0, np.uint32)
    return np.concatenate(parts).astype(np.uint32)


def density_bitmap(rng, universe, density):
    """Bernoulli(density) subset of [0, universe)."""
    return np.flatnonzero(rng.random(universe) < density).astype(np.uint32)
