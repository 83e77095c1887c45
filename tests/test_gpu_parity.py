"""GPU parity tests: the CUDA path (through the C ABI) vs the unmodified reference
(oracle/_ref), the plain-C oracle and the committed golden values — byte-exact on the
portable serialization (keys, container TYPES, cardinalities, payloads).

Mirrors /root/reference/tests/realdata_unit.c:365-756 (compare_intersections / unions / xors /
andnots / wide_unions on the real-data sets, run and no-run twins) and the cell-level tests of
/root/reference/tests/mixed_container_unit.c (every A/B/R pairing).
"""
import numpy as np
import pytest

from helpers import DATASETS, OPS, check_result_bitmap, no_run_twins, sha_concat, synth_blobs

pytestmark = pytest.mark.gpu


def _successive(n):
    return np.arange(n - 1, dtype=np.uint32), np.arange(1, n, dtype=np.uint32)


@pytest.mark.parametrize("ds", DATASETS)
def test_realdata_successive_pairs(rb, R, O, golden, ds):
    """config[0]/[1]: pairwise op of successive real-data bitmaps, byte-exact vs reference."""
    blobs = rb.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    ia, ib = _successive(len(blobs))
    for op in OPS:
        res = S.batch(op, S, ia, ib)
        cards = res.cardinalities()
        outs = res.download_all()
        got = [o.serialize() for o in outs]
        g = golden[ds]["run_optimized"][op]
        assert int(cards.sum()) == g["sum_card"], (ds, op)
        assert sha_concat(got) == g["sha256"], (ds, op)
        # spot-check through the reference + oracle objects too
        for k in (0, 1, len(outs) // 2, len(outs) - 1):
            exp = R.op_bytes(op, blobs[k], blobs[k + 1])
            check_result_bitmap(R, outs[k], exp, f"{ds} {op} pair {k}")
            assert O.op_bytes(op, blobs[k], blobs[k + 1]) == exp
            assert int(cards[k]) == outs[k].cardinality()
        for o in outs:
            o.free()
        res.free()
    S.free()


@pytest.mark.parametrize("ds", ["census1881", "weather_sept_85", "wikileaks-noquotes"])
def test_realdata_no_run_twins(rb, R, golden, ds):
    """Same sweep on the twins without run containers (tests/realdata_unit.c:796-806)."""
    blobs = no_run_twins(R, rb.load_realdata(ds))
    S = rb.DeviceSet.from_serialized(blobs)
    ia, ib = _successive(len(blobs))
    for op in OPS:
        res = S.batch(op, S, ia, ib)
        outs = res.download_all()
        g = golden[ds]["no_runs"][op]
        assert sum(o.cardinality() for o in outs) == g["sum_card"], (ds, op)
        assert sha_concat([o.serialize() for o in outs]) == g["sha256"], (ds, op)
        res.free()
    S.free()


@pytest.mark.parametrize("ds", DATASETS)
def test_realdata_and_cardinality(rb, R, golden, ds):
    blobs = rb.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    ia, ib = _successive(len(blobs))
    c = S.and_cardinality(S, ia, ib)
    assert int(c.sum()) == golden[ds]["run_optimized"]["and_cardinality"]
    for k in (0, 7, 100, 198):
        ra, rbm = R.deserialize(blobs[k]), R.deserialize(blobs[k + 1])
        assert int(c[k]) == int(R.L.roaring_bitmap_and_cardinality(ra, rbm))
        R.free(ra), R.free(rbm)
    S.free()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_synthetic_all_cells(rb, R, O, seed):
    """Seeded container mixes hitting all 9 type pairings, full / near-full / edge-4096 containers."""
    blobs = synth_blobs(R, seed, 96)
    rng = np.random.default_rng(seed + 100)
    ia = rng.integers(0, len(blobs), 700).astype(np.uint32)
    ib = rng.integers(0, len(blobs), 700).astype(np.uint32)
    ia[:8] = ib[:8]  # x op x
    S = rb.DeviceSet.from_serialized(blobs)
    for op in OPS:
        res = S.batch(op, S, ia, ib)
        outs = res.download_all()
        for k, o in enumerate(outs):
            exp = R.op_bytes(op, blobs[ia[k]], blobs[ib[k]])
            check_result_bitmap(R, o, exp, f"seed {seed} {op} pair {k} ({ia[k]},{ib[k]})")
        # the oracle agrees with the reference on a sample (pins the oracle on the GPU box too)
        for k in range(0, 700, 50):
            assert O.op_bytes(op, blobs[ia[k]], blobs[ib[k]]) == outs[k].serialize()
        res.free()
    c = S.and_cardinality(S, ia, ib)
    for k in range(0, 700, 7):
        assert int(c[k]) == O.and_cardinality(blobs[ia[k]], blobs[ib[k]])
    S.free()


def test_two_sets_and_chaining(rb, R):
    """A and B different sets; results of one op feed the next op without leaving the device."""
    a = synth_blobs(R, 11, 40)
    b = synth_blobs(R, 12, 40)
    A, B = rb.DeviceSet.from_serialized(a), rb.DeviceSet.from_serialized(b)
    idx = np.arange(40, dtype=np.uint32)
    r1 = A.batch("or", B, idx, idx)          # a | b
    r2 = r1.batch("andnot", A, idx, idx)     # (a | b) \ a
    r3 = r2.batch("xor", B, idx, idx[::-1].copy())
    outs = r3.download_all()
    for k in range(40):
        e1 = R.op_bytes("or", a[k], b[k])
        e2 = R.op_bytes("andnot", e1, a[k])
        e3 = R.op_bytes("xor", e2, b[39 - k])
        check_result_bitmap(R, outs[k], e3, f"chain {k}")


def test_empty_and_degenerate(rb, R):
    empty = R.serialize(R.L.roaring_bitmap_create_with_capacity(0))
    one = R.serialize(R.from_values(np.array([5], dtype=np.uint32)))
    full = R.serialize(R.from_values(np.arange(0, 3 * 65536, dtype=np.uint32)))
    blobs = [empty, one, full, empty]
    S = rb.DeviceSet.from_serialized(blobs)
    ia = np.array([0, 0, 1, 2, 2, 3, 1, 2], dtype=np.uint32)
    ib = np.array([0, 1, 0, 2, 1, 2, 1, 0], dtype=np.uint32)
    for op in OPS:
        res = S.batch(op, S, ia, ib)
        for k, o in enumerate(res.download_all()):
            check_result_bitmap(R, o, R.op_bytes(op, blobs[ia[k]], blobs[ib[k]]), f"{op} {k}")
    # zero pairs
    res = S.batch("or", S, np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    assert len(res) == 0
    assert S.and_cardinality(S, ia, ib).tolist() == [0, 0, 0, 3 * 65536, 1, 0, 1, 0]


@pytest.mark.parametrize("ds", DATASETS)
def test_realdata_or_many(rb, R, golden, ds):
    """config[2]-style N-way union on the real-data sets vs roaring_bitmap_or_many."""
    blobs = rb.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    res = S.or_many()
    out = res.download(0)
    g = golden[ds]["run_optimized"]["or_many"]
    assert out.cardinality() == g["card"]
    assert int(res.cardinalities()[0]) == g["card"]
    assert sha_concat([out.serialize()]) == g["sha256"]
    check_result_bitmap(R, out, R.many_bytes("or_many", blobs), f"{ds} or_many")
    # no-run twins
    nr = no_run_twins(R, blobs)
    S2 = rb.DeviceSet.from_serialized(nr)
    out2 = S2.or_many().download(0)
    assert sha_concat([out2.serialize()]) == golden[ds]["no_runs"]["or_many"]["sha256"]


@pytest.mark.parametrize("seed", [5, 6, 7, 8])
def test_synthetic_or_many_state_machine(rb, R, O, seed):
    """Random subsets / orders with full, near-full and half containers: exercises the
    reference's full-container state machine incl. the ordered replay."""
    profiles = ["full", "nearfull", "halves", "dense", "bitset", "array", "tiny", "longruns",
                "shortruns", "ends"]
    blobs = synth_blobs(R, seed, 60, key_space=6, max_keys=7, profiles=profiles)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(seed)
    for trial in range(60):
        n = int(rng.integers(0, 12))
        idx = rng.integers(0, len(blobs), n).astype(np.uint32)
        res = S.or_many(idx) if n else S.or_many(np.zeros(0, np.uint32))
        out = res.download(0)
        sub = [blobs[i] for i in idx]
        exp = R.many_bytes("or_many", sub)
        assert O.many_bytes("or_many", sub) == exp
        check_result_bitmap(R, out, exp, f"seed {seed} trial {trial} idx {idx.tolist()}")


def test_or_many_t4_named_cases(rb, R):
    """The probes of SURVEY.md §8a-T4: even/odd halves, rest+small, full-run first."""
    even = np.arange(0, 65536, 2, dtype=np.uint32)
    odd = np.arange(1, 65536, 2, dtype=np.uint32)
    sm = np.array([3, 9, 77], dtype=np.uint32)
    rest = np.setdiff1d(np.arange(65536, dtype=np.uint32), sm)
    fullv = np.arange(65536, dtype=np.uint32)
    mk = lambda v, ro: R.serialize(R.from_values(v, run_optimize=ro))
    cases = {
        "even,odd": [mk(even, False), mk(odd, False)],
        "rest,sm,even": [mk(rest, False), mk(sm, False), mk(even, False)],
        "rest,sm": [mk(rest, False), mk(sm, False)],
        "fullrun,sm": [mk(fullv, True), mk(sm, False)],
        "fullrun,sm,even": [mk(fullv, True), mk(sm, False), mk(even, False)],
        "sm,fullrun": [mk(sm, False), mk(fullv, True)],
        "even,sm,fullrun,odd": [mk(even, False), mk(sm, False), mk(fullv, True), mk(odd, False)],
        "fullbitset,even": [mk(fullv, False), mk(even, False)],
        "sm,fullbitset,even,odd": [mk(sm, False), mk(fullv, False), mk(even, False), mk(odd, False)],
        "even,even,odd": [mk(even, False), mk(even, False), mk(odd, False)],
        "sm,even,odd,even": [mk(sm, False), mk(even, False), mk(odd, False), mk(even, False)],
        "single run": [mk(np.arange(100, 9000, dtype=np.uint32), True)],
        "single": [mk(even, False)],
    }
    for name, blobs in cases.items():
        S = rb.DeviceSet.from_serialized(blobs)
        out = S.or_many().download(0)
        check_result_bitmap(R, out, R.many_bytes("or_many", blobs), name)


def test_dropin_symbols(rb, R):
    """roaring_bitmap_{and,or,xor,andnot,or_many,and_cardinality,...} on host bitmaps."""
    blobs = rb.load_realdata("weather_sept_85")[:6] + synth_blobs(R, 21, 6)
    bms = [rb.Bitmap.deserialize(b) for b in blobs]
    for i in range(len(bms) - 1):
        a, b = bms[i], bms[i + 1]
        for op, f in (("and", lambda x, y: x & y), ("or", lambda x, y: x | y),
                      ("xor", lambda x, y: x ^ y), ("andnot", lambda x, y: x - y)):
            check_result_bitmap(R, f(a, b), R.op_bytes(op, blobs[i], blobs[i + 1]), f"dropin {op} {i}")
        ra, rbm = R.deserialize(blobs[i]), R.deserialize(blobs[i + 1])
        assert a.and_cardinality(b) == int(R.L.roaring_bitmap_and_cardinality(ra, rbm))
        assert a.or_cardinality(b) == int(R.L.roaring_bitmap_or_cardinality(ra, rbm))
        assert a.xor_cardinality(b) == int(R.L.roaring_bitmap_xor_cardinality(ra, rbm))
        assert a.andnot_cardinality(b) == int(R.L.roaring_bitmap_andnot_cardinality(ra, rbm))
        assert a.jaccard_index(b) == R.L.roaring_bitmap_jaccard_index(ra, rbm)
        assert a.intersect(b) == bool(R.L.roaring_bitmap_intersect(ra, rbm))
        R.free(ra), R.free(rbm)
    check_result_bitmap(R, rb.or_many(bms), R.many_bytes("or_many", blobs), "dropin or_many")
    # the reference's own functions accept what we return, and free it
    r = bms[0] | bms[1]
    r.own = False
    R.free(r.ptr)


def test_batch_op_host(rb, R):
    blobs = rb.load_realdata("census1881")[:64]
    bms = [rb.Bitmap.deserialize(b) for b in blobs]
    for op in OPS:
        outs = rb.batch_op_host(op, bms[:-1], bms[1:])
        for k, o in enumerate(outs):
            check_result_bitmap(R, o, R.op_bytes(op, blobs[k], blobs[k + 1]), f"host {op} {k}")


def test_cow_flag_propagates(rb, R):
    a = R.from_values(np.arange(0, 100000, 3, dtype=np.uint32))
    b = R.from_values(np.arange(0, 100000, 5, dtype=np.uint32))
    R.L.roaring_bitmap_set_copy_on_write(a, True)
    A, B = rb.Bitmap(a, own=False), rb.Bitmap(b, own=False)
    r = A | B
    import ctypes
    flags = ctypes.cast(r.ptr, ctypes.POINTER(ctypes.c_uint8 * 40)).contents[32]
    assert flags & 1  # src/roaring.c:890
    exp = R.op(
        "or", a, b)
    assert r.serialize() == R.serialize(exp)
    R.free(exp), R.free(a), R.free(b)


def test_large_properties_bitset_heavy(rb):
    """config[3] shape at full size (bitset-heavy, 2^20 universe): size-independent properties
    |a&b| + |a^b| + ... inclusion-exclusion identities and idempotence, no reference needed."""
    rng = np.random.default_rng(99)
    n = 256
    words = rng.integers(0, 2 ** 63, size=(n, 16, 1024), dtype=np.int64).view(np.uint64)
    blobs = []
    for i in range(n):
        # hand-built portable bitmap: 16 bitset containers (cookie 12346, no runs)
        cards = np.array([int(np.unpackbits(words[i, k].view(np.uint8)).sum()) for k in range(16)])
        hdr = np.zeros(2 + 16 + 16, dtype=np.uint32)
        hdr[0], hdr[1] = 12346, 16
        kc = np.zeros(32, dtype=np.uint16)
        kc[0::2] = np.arange(16)
        kc[1::2] = (cards - 1).astype(np.uint16)
        hdr[2:18] = kc.view(np.uint32)
        hdr[18:34] = (8 + 8 * 16 + 8192 * np.arange(16)).astype(np.uint32)
        blobs.append(hdr.tobytes() + words[i].tobytes())
    S = rb.DeviceSet.from_serialized(blobs)
    ia = np.arange(0, n, 2, dtype=np.uint32)
    ib = ia + 1
    c_and = S.and_cardinality(S, ia, ib)
    card = S.cardinalities()
    r_and, r_or, r_xor, r_andn = (S.batch(op, S, ia, ib) for op in OPS)
    ca, co, cx, cn = (r.cardinalities() for r in (r_and, r_or, r_xor, r_andn))
    assert (ca == c_and).all()
    assert (co == card[ia] + card[ib] - c_and).all()
    assert (cx == co - ca).all()
    assert (cn == card[ia] - c_and).all()
    # numpy cross-check of one pair
    exp = int(np.unpackbits((words[0] & words[1]).view(np.uint8)).sum())
    assert int(c_and[0]) == exp
    # idempotence: (a|b)|b == a|b ; (a&b)&b == a&b  (byte-exact)
    k = np.arange(len(ia), dtype=np.uint32)
    again = r_or.batch("or", S, k, ib)
    assert [x.serialize() for x in again.download_all()] == [x.serialize() for x in r_or.download_all()]
    again = r_and.batch("and", S, k, ib)
    assert [x.serialize() for x in again.download_all()] == [x.serialize() for x in r_and.download_all()]


def test_streaming_download_matches(rb, R):
    """rb200_download_begin/next/end yields the same bitmaps, in order, as download_all."""
    blobs = rb.load_realdata("weather_sept_85")[:60]
    S = rb.DeviceSet.from_serialized(blobs)
    i, j = np.triu_indices(len(blobs), 1)
    res = S.batch("xor", S, i.astype(np.uint32), j.astype(np.uint32))
    ref_all = [o.serialize() for o in res.download_all()]
    got = []
    for arr, n in res.download_stream(97):
        assert 0 < n <= 97
        for k in range(n):
            b = rb.Bitmap(arr[k])
            ok, why = R.validate(b.ptr)
            assert ok, why
            got.append(b.serialize())
            b.own = False
        rb.DeviceSet.free_raw(arr, n)
    assert got == ref_all
    for k in (0, 500, len(got) - 1):
        assert got[k] == R.op_bytes("xor", blobs[i[k]], blobs[j[k]])
    # empty set streams nothing
    e = S.batch("and", S, np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    assert list(e.download_stream()) == []


def test_relations_equals_subset(rb, R):
    """roaring_bitmap_equals / is_subset / is_strict_subset (roaring.h:899-912), batched and as
    drop-in symbols, vs the reference — on pairs built to be equal, nested, and unrelated."""
    blobs = synth_blobs(R, 77, 40, key_space=5, max_keys=6)
    rs = [R.deserialize(b) for b in blobs]
    # derived bitmaps: unions (supersets), intersections (subsets), copies (equal, other encoding)
    extra = []
    for k in range(0, 38, 2):
        extra.append(R.op_bytes("or", blobs[k], blobs[k + 1]))
        extra.append(R.op_bytes("and", blobs[k], blobs[k + 1]))
    twins = []
    for b in blobs[:20]:
        r = R.deserialize(b)
        R.L.roaring_bitmap_remove_run_compression(r)
        twins.append(R.serialize(r))
        R.free(r)
    allb = blobs + extra + twins
    ra = [R.deserialize(b) for b in allb]
    S = rb.DeviceSet.from_serialized(allb)
    rng = np.random.default_rng(5)
    ia = np.concatenate([rng.integers(0, len(allb), 600), np.arange(20), np.arange(0, 38, 2), np.arange(40, 78)]).astype(np.uint32)
    ib = np.concatenate([rng.integers(0, len(allb), 600), np.arange(78, 98), np.arange(40, 78, 2), np.repeat(np.arange(0, 38, 2), 2)]).astype(np.uint32)
    got = S.relations(S, ia, ib)
    seen = set()
    for k in range(len(ia)):
        a, b = ra[ia[k]], ra[ib[k]]
        exp = (1 if R.L.roaring_bitmap_equals(a, b) else 0) | (2 if R.L.roaring_bitmap_is_subset(a, b) else 0) | \
              (4 if R.L.roaring_bitmap_is_strict_subset(a, b) else 0)
        assert got[k] == exp, (k, int(ia[k]), int(ib[k]))
        seen.add(exp)
    assert {0, 3, 6} <= seen
    ha, hb = rb.Bitmap.deserialize(allb[0]), rb.Bitmap.deserialize(allb[78])
    assert ha.equals(hb) and ha.is_subset(hb) and not ha.is_strict_subset(hb)
    hu = rb.Bitmap.deserialize(extra[0])
    assert ha.is_subset(hu) == bool(R.L.roaring_bitmap_is_subset(ra[0], ra[40]))
    assert ha.is_strict_subset(hu) == bool(R.L.roaring_bitmap_is_strict_subset(ra[0], ra[40]))
    for x in rs + ra:
        R.free(x)
