"""The C workload generators (croaring_b200/csrc/workgen.c) against the reference: the bytes they
emit are exactly roaring_bitmap_of_ptr + run_optimize + portable_serialize of the same values,
and the value streams follow the PCG32 definitions of SURVEY.md §8(d) rows 3-5."""
import math

import numpy as np

from croaring_b200 import workloads as wl


def pcg32(state, inc):
    M = (1 << 64) - 1
    while True:
        old = state
        state = (old * 6364136223846793005 + inc) & M
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        yield ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF


def zipf_values(b, U, n, density_draw=False):
    g = pcg32(0x853c49e6748fea9b ^ b, 0xda3e39cb94b95bdb)
    next(g)                                      # warm-up step (see csrc/workgen.c)
    if density_draw:
        u0 = (next(g) + 1) / 4294967296.0
        n = max(1, int(round(0.001 * math.pow(300.0, u0) * U)))
    n = min(n, U)
    seen = set()
    l2 = math.log2(U)
    while len(seen) < n:
        seen.add(zipf_value(next(g), l2, U))
    return np.array(sorted(seen), dtype=np.uint32)


T256 = [math.pow(2.0, j / 256.0) for j in range(257)]


def zipf_value(r, log2u, U):
    """csrc/workgen.c zipf_value(), operation by operation (IEEE doubles, no fused multiply-add)."""
    u = (float(r) + 1.0) * (1.0 / 4294967296.0)
    x = u * log2u
    xi = math.floor(x)
    f = x - xi
    fj = math.floor(f * 256.0)
    t = (f - fj * (1.0 / 256.0)) * 0.6931471805599453
    p = 1.0 + t * (1.0 + t * (0.5 + t * ((1.0 / 6.0) + t * (1.0 / 24.0))))
    y = T256[int(fj)] * p * float(1 << int(xi))
    v = math.floor(y) - 1.0
    if v < 0.0:
        v = 0.0
    return min(int(v), U - 1)


def test_zipf_matches_definition_and_reference(R):
    # ((1 << 28) + 12345: large universes take the batched / partitioned path of the generator)
    for (U, n, dd, ro) in [(300000, 20000, False, True), (1 << 20, 3000, False, True), ((1 << 28) + 12345, 150000, False, True),
                           (200000, None, True, True), (300000, 20000, False, False)]:
        A = wl.zipf_arena(3, U, n, b0=5, density_draw=dd, run_optimize=ro, threads=2)
        for i in range(3):
            vals = zipf_values(5 + i, U, n or 0, dd)
            assert int(A.cards[i]) == len(vals)
            r = R.from_values(vals, run_optimize=ro)
            assert A.blob(i) == R.serialize(r), (U, n, dd, ro, i)
            R.free(r)
        A.free()


def test_zipf_saturates_low_keys(R):
    """Low keys of a dense Zipf bitmap are full containers (the or_many state machine T4 needs them)."""
    A = wl.zipf_arena(1, wl.zipf_universe(10 ** 7, 0.3), 10 ** 7, threads=1)   # one config-3 bitmap, d = 0.3
    r = R.deserialize(A.blob(0))
    assert R.card(r) == 10 ** 7
    vals = R.to_array(r)
    assert np.array_equal(vals[:65536], np.arange(65536, dtype=np.uint32))
    assert A.blob(0)[:2] == b"\x3b\x30"      # SERIAL_COOKIE: the saturated key is a run container
    R.free(r)
    A.free()


def test_dense_is_the_global_pcg_stream(R):
    nk = 2
    A = wl.dense_arena(3, n_keys=nk, i0=1, threads=2)
    g = pcg32(0x853c49e6748fea9b, 0xda3e39cb94b95bdb)
    per = nk << 16
    for _ in range(per):          # bitmap 0 of the stream is skipped (i0 = 1)
        next(g)
    for i in range(3):
        bits = np.fromiter((next(g) & 1 for _ in range(per)), dtype=np.uint8, count=per)
        vals = np.flatnonzero(bits).astype(np.uint32)
        r = R.from_values(vals, run_optimize=False)
        assert A.blob(i) == R.serialize(r), i
        R.free(r)
    A.free()


def test_cached_arena_roundtrip(tmp_path):
    """The /dev/shm memoisation of the generators returns the very bytes the generator emitted."""
    build = lambda: wl.zipf_arena(4, 10 ** 6, None, b0=3, density_draw=True, threads=2)
    a = wl.cached_arena("t", build, cache_dir=str(tmp_path))
    first, cards = a.blobs(), [int(x) for x in a.cards]
    b = wl.cached_arena("t", build, cache_dir=str(tmp_path))
    assert type(b).__name__ == "MappedArena" and b.blobs() == first and [int(x) for x in b.cards] == cards
    assert wl.cached_arena("t", build, cache_dir="").blobs() == first      # caching disabled: rebuilt, same bytes
