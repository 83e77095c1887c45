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
