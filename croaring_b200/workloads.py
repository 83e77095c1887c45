"""Synthetic workload generators of BASELINE.json configs[2..4], numpy only.

They emit portable-serialized bitmaps directly (the format of
/root/reference/src/roaring_array.c:469-531), so neither the reference nor the oracle is needed
to build inputs; parity tests feed the same bytes to both sides.
"""
import numpy as np

SERIAL_COOKIE_NO_RUN = 12346
SERIAL_COOKIE = 12347


def serialize_containers(keys, conts):
    """Portable serialization of containers given as (type, payload ndarray) per key.

    type 'b': 1024 x u64 words; 'a': sorted u16 values; 'r': (n,2) u16 (start, length-1).
    """
    n = len(keys)
    hasrun = any(t == "r" for t, _ in conts)
    cards, sizes = [], []
    for t, p in conts:
        if t == "b":
            cards.append(int(np.unpackbits(p.view(np.uint8)).sum()))
            sizes.append(8192)
        elif t == "a":
            cards.append(len(p))
            sizes.append(2 * len(p))
        else:
            cards.append(int(p[:, 1].astype(np.int64).sum()) + len(p))
            sizes.append(2 + 4 * len(p))
    parts = []
    if hasrun:
        parts.append(np.array([SERIAL_COOKIE | ((n - 1) << 16)], dtype=np.uint32).tobytes())
        flags = np.zeros((n + 7) // 8, dtype=np.uint8)
        for i, (t, _) in enumerate(conts):
            if t == "r":
                flags[i // 8] |= 1 << (i % 8)
        parts.append(flags.tobytes())
        hdr = 4 + len(flags) + (4 * n if n < 4 else 8 * n)
    else:
        parts.append(np.array([SERIAL_COOKIE_NO_RUN, n], dtype=np.uint32).tobytes())
        hdr = 8 + 8 * n
    kc = np.zeros(2 * n, dtype=np.uint16)
    kc[0::2] = np.asarray(keys, dtype=np.uint16)
    kc[1::2] = (np.asarray(cards, dtype=np.int64) - 1).astype(np.uint16)
    parts.append(kc.tobytes())
    if (not hasrun) or n >= 4:
        offs = hdr + np.concatenate([[0], np.cumsum(sizes)[:-1]]) if n else np.zeros(0)
        parts.append(np.asarray(offs, dtype=np.uint32).tobytes())
    for t, p in conts:
        if t == "b":
            parts.append(np.ascontiguousarray(p, dtype=np.uint64).tobytes())
        elif t == "a":
            parts.append(np.ascontiguousarray(p, dtype=np.uint16).tobytes())
        else:
            parts.append(np.array([len(p)], dtype=np.uint16).tobytes())
            parts.append(np.ascontiguousarray(p, dtype=np.uint16).tobytes())
    return b"".join(parts)


def bitset_heavy_blobs(n, seed=0, keys=16):
    """configs[3]: bitmaps over universe keys*2^16 with density 0.5 -> `keys` bitset containers."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        w = rng.integers(0, 2 ** 63, size=(keys, 1024), dtype=np.int64).view(np.uint64)
        w ^= rng.integers(0, 2, size=(keys, 1024), dtype=np.int64).view(np.uint64) << np.uint64(63)
        out.append(serialize_containers(list(range(keys)), [("b", w[k]) for k in range(keys)]))
    return out


def values_to_blob(vals):
    """Sorted unique uint32 values -> portable bitmap with array / bitset containers only
    (array iff card <= 4096, as roaring_bitmap_of_ptr without run_optimize)."""
    vals = np.asarray(vals, dtype=np.uint32)
    hi = (vals >> np.uint32(16)).astype(np.uint32)
    keys, starts = np.unique(hi, return_index=True)
    ends = np.append(starts[1:], len(vals))
    conts = []
    for s, e in zip(starts, ends):
        low = (vals[s:e] & np.uint32(0xFFFF)).astype(np.uint16)
        if e - s <= 4096:
            conts.append(("a", low))
        else:
            bits = np.zeros(65536, dtype=np.uint8)
            bits[low] = 1
            conts.append(("b", np.packbits(bits, bitorder="little").view(np.uint64)))
    return serialize_containers(keys.tolist(), conts)


# ------------------------------------------------------------------------------------------
# Zipfian many-way-union inputs (BASELINE.json configs[2] and [4]).
#
# SURVEY.md defines bitmap b as "draw v = floor(U^u) - 1, u ~ U(0,1], until n distinct values":
# P(v) ∝ 1/(v+1) (Zipf s=1).  Drawing ~n*ln(U) samples per bitmap is far too slow for a bench
# that must finish in minutes, so we generate the Poissonised form of the same process: value v
# is present independently with probability 1 - exp(-c/(v+1)), c chosen so that the expected
# cardinality is n.  Per 2^16-value container the inclusion density is (nearly) constant, so a
# container is filled either by thresholding random bits (density >= 1/16, quantised to 1/16)
# or by geometric gap sampling (sparse).  Low keys saturate (full containers -> exercises the
# or_many full-container state machine), high keys are sparse arrays.
def _zipf_key_density(n_keys, n_values):
    mid = (np.arange(n_keys, dtype=np.float64) * 65536.0 + 32768.0)
    lo, hi = 1e-3, 1e12
    for _ in range(200):
        c = np.sqrt(lo * hi)
        tot = (1.0 - np.exp(-c / mid)).sum() * 65536.0
        if tot > n_values:
            hi = c
        else:
            lo = c
    return 1.0 - np.exp(-np.sqrt(lo * hi) / mid)


def _random_words(rng, n_words, sixteenths):
    """n_words u64 with each bit set with probability sixteenths/16 (1..16)."""
    if sixteenths >= 16:
        return np.full(n_words, np.uint64(0xFFFFFFFFFFFFFFFF))
    r = rng.integers(0, 2 ** 63, size=(4, n_words), dtype=np.int64).view(np.uint64)
    r = r ^ (rng.integers(0, 2, size=(4, n_words), dtype=np.int64).view(np.uint64) << np.uint64(63))
    acc = np.zeros(n_words, dtype=np.uint64)          # probability 0
    for k in range(4):                                 # binary expansion, LSB first
        acc = (acc | r[k]) if (sixteenths >> k) & 1 else (acc & r[k])
    return acc


def zipf_bitmap_blob(rng, n_values, density, run_optimize_full=True):
    """One portable bitmap with ~n_values Zipf-distributed values over universe n_values/density
    (clamped to 2^32).  Saturated containers are emitted as the full run [0,65535]."""
    universe = min(2 ** 32, int(np.ceil(n_values / density)))
    n_keys = (universe + 65535) // 65536
    dens = _zipf_key_density(n_keys, n_values)
    keys, conts = [], []
    dense = np.flatnonzero(dens >= 1.0 / 16)
    for k in dense:
        q = int(min(16, max(1, round(dens[k] * 16))))
        if q >= 16 and run_optimize_full:
            conts.append(("r", np.array([[0, 65535]], dtype=np.uint16)))
        else:
            w = _random_words(rng, 1024, q)
            card = int(np.unpackbits(w.view(np.uint8)).sum())
            if card <= 4096:
                bits = np.unpackbits(w.view(np.uint8), bitorder="little")
                conts.append(("a", np.flatnonzero(bits).astype(np.uint16)))
            else:
                conts.append(("b", w))
        keys.append(int(k))
    # sparse tail: geometric gaps over the concatenated remaining universe, piecewise by key
    sparse = np.flatnonzero(dens < 1.0 / 16)
    if len(sparse):
        # one gap stream per block of keys with similar density keeps this vectorised
        for blk in np.array_split(sparse, max(1, len(sparse) // 64)):
            if len(blk) == 0:
                continue
            d = float(dens[blk].mean())
            span = len(blk) * 65536
            m = int(span * d * 1.3 + 64)
            pos = np.cumsum(rng.geometric(d, size=m)) - 1
            pos = pos[pos < span]
            kk = pos >> 16
            low = (pos & 0xFFFF).astype(np.uint16)
            ks, starts = np.unique(kk, return_index=True)
            ends = np.append(starts[1:], len(pos))
            for kidx, s, e in zip(ks, starts, ends):
                if e - s > 4096:                      # cannot happen for d < 1/16 (mean 4096) but be safe
                    bits = np.zeros(65536, dtype=np.uint8)
                    bits[low[s:e]] = 1
                    conts.append(("b", np.packbits(bits, bitorder="little").view(np.uint64)))
                else:
                    conts.append(("a", low[s:e]))
                keys.append(int(blk[int(kidx)]))
    order = np.argsort(keys, kind="stable")
    return serialize_containers([keys[i] for i in order], [conts[i] for i in order])


def zipf_blobs(n_bitmaps, n_values, density, seed=0):
    return [zipf_bitmap_blob(np.random.default_rng(seed * 100003 + b), n_values, density)
            for b in range(n_bitmaps)]
