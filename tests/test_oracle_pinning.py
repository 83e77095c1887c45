"""CPU tests (-m "not gpu"): pin the plain-C oracle (oracle/roaring_oracle.c) against
(a) the committed golden values generated from the reference and (b) the unmodified reference
itself (oracle/_ref/libroaring_ref.so), byte-for-byte on the portable serialization."""
import hashlib

import numpy as np
import pytest

import croaring_b200.datasets as dsm
from helpers import DATASETS, OPS, no_run_twins, sha_concat, synth_blobs


@pytest.mark.parametrize("ds", DATASETS)
def test_oracle_vs_golden_realdata(O, golden, ds):
    blobs = dsm.load_realdata(ds)
    assert len(blobs) == golden[ds]["n"] == 200
    assert sum(map(len, blobs)) == golden[ds]["portable_bytes"]
    g = golden[ds]["run_optimized"]
    for op in OPS:
        outs = [O.op_bytes(op, blobs[i], blobs[i + 1]) for i in range(199)]
        assert sum(O.cardinality(b) for b in outs) == g[op]["sum_card"], (ds, op)
        assert sha_concat(outs) == g[op]["sha256"], (ds, op)
    assert sum(O.and_cardinality(blobs[i], blobs[i + 1]) for i in range(199)) == g["and_cardinality"]
    for name in ("or_many", "xor_many"):
        out = O.many_bytes(name, blobs)
        assert O.cardinality(out) == g[name]["card"]
        assert hashlib.sha256(out).hexdigest() == g[name]["sha256"], (ds, name)


@pytest.mark.parametrize("ds", ["census1881", "wikileaks-noquotes", "weather_sept_85"])
def test_oracle_vs_golden_no_runs(O, R, golden, ds):
    blobs = no_run_twins(R, dsm.load_realdata(ds))
    g = golden[ds]["no_runs"]
    for op in OPS:
        outs = [O.op_bytes(op, blobs[i], blobs[i + 1]) for i in range(199)]
        assert sha_concat(outs) == g[op]["sha256"], (ds, op)
    assert hashlib.sha256(O.many_bytes("or_many", blobs)).hexdigest() == g["or_many"]["sha256"]
    assert hashlib.sha256(O.many_bytes("xor_many", blobs)).hexdigest() == g["xor_many"]["sha256"]


def test_survey_golden_sum_cards(golden):
    """The checksums measured in SURVEY.md §6 / BASELINE.md §2."""
    exp = {"census1881": (23, 2007691, 2007668, 1003836, 988653, 973455),
           "weather_sept_85": (642019, 24729002, 24086983, 11960876, 1015367, 526545),
           "wikileaks-noquotes": (3327, 541893, 538566, 271605, 242540, 212267)}
    for ds, (a, o, x, n, om, xm) in exp.items():
        g = golden[ds]["run_optimized"]
        assert (g["and"]["sum_card"], g["or"]["sum_card"], g["xor"]["sum_card"],
                g["andnot"]["sum_card"], g["or_many"]["card"], g["xor_many"]["card"]) == (a, o, x, n, om, xm)


@pytest.mark.parametrize("seed", [1234, 99])
def test_oracle_vs_reference_synthetic(O, R, seed):
    blobs = synth_blobs(R, seed, 100)
    rng = np.random.default_rng(seed)
    for _ in range(400):
        i, j = rng.integers(0, len(blobs), 2)
        for op in OPS:
            assert O.op_bytes(op, blobs[i], blobs[j]) == R.op_bytes(op, blobs[i], blobs[j]), (op, i, j)
        ra, rb_ = R.deserialize(blobs[i]), R.deserialize(blobs[j])
        assert O.and_cardinality(blobs[i], blobs[j]) == int(R.L.roaring_bitmap_and_cardinality(ra, rb_))
        R.free(ra), R.free(rb_)
    for _ in range(200):
        idx = rng.integers(0, len(blobs), int(rng.integers(0, 9)))
        sub = [blobs[k] for k in idx]
        for name in ("or_many", "xor_many"):
            assert O.many_bytes(name, sub) == R.many_bytes(name, sub), (name, idx.tolist())


def test_reference_known_answers(R, O):
    """Known answers of the reference's own tests: tests/toplevel_unit.c:1730-1806
    (array x array -> 2*34 ; bitset x bitset -> 26666)."""
    a = np.arange(0, 100, dtype=np.uint32) * 2
    b = np.arange(0, 100, dtype=np.uint32) * 3
    r1, r2 = R.from_values(a, False), R.from_values(b, False)
    exp = len(np.intersect1d(a, b))
    assert O.and_cardinality(R.serialize(r1), R.serialize(r2)) == exp == 34
    x1 = R.from_values(np.arange(0, 20000, dtype=np.uint32) * 2, False)   # toplevel_unit.c:1781
    x2 = R.from_values(np.arange(0, 20000, dtype=np.uint32) * 3, False)
    inter = len(np.intersect1d(np.arange(20000) * 2, np.arange(20000) * 3))
    out = O.op_bytes("and", R.serialize(x1), R.serialize(x2))
    assert O.cardinality(out) == inter == 6667
    assert out == R.op_bytes("and", R.serialize(x1), R.serialize(x2))
    for r in (r1, r2, x1, x2):
        R.free(r)


@pytest.mark.parametrize("seed", [7, 4242])
def test_oracle_lazy_api_and_heap_vs_reference(O, R, seed):
    """The public lazy API (roaring.h:932-977) folded left to right then repaired, for both values
    of `bitsetconversion`, and roaring_bitmap_or_many_heap (roaring_priority_queue.c:200)."""
    blobs = synth_blobs(R, seed, 80)
    rng = np.random.default_rng(seed)
    for _ in range(150):
        idx = rng.integers(0, len(blobs), int(rng.integers(1, 8)))
        sub = [blobs[k] for k in idx]
        for conv in (False, True):
            assert O.lazy_fold_bytes("or", conv, sub) == R.lazy_fold_bytes("or", conv, sub), (conv, idx.tolist())
        assert O.lazy_fold_bytes("xor", False, sub) == R.lazy_fold_bytes("xor", False, sub), idx.tolist()
        assert O.or_many_heap_bytes(sub) == R.many_bytes("or_many_heap", sub), idx.tolist()
    assert O.or_many_heap_bytes([]) == R.many_bytes("or_many_heap", [])


@pytest.mark.parametrize("ds", ["census1881", "weather_sept_85", "wikileaks-noquotes", "uscensus2000"])
def test_oracle_heap_vs_reference_realdata(O, R, ds):
    blobs = dsm.load_realdata(ds)
    assert O.or_many_heap_bytes(blobs) == R.many_bytes("or_many_heap", blobs)
    assert O.or_many_heap_bytes(blobs[:37]) == R.many_bytes("or_many_heap", blobs[:37])
    assert O.lazy_fold_bytes("or", False, blobs[:50]) == R.lazy_fold_bytes("or", False, blobs[:50])


FLIP_RANGES = [(0, 1), (5, 6), (5, 7), (0, 65536), (0, 65537), (65535, 65537), (100, 70000), (65536, 131072),
               (1 << 16, (5 << 16) + 17), (3, 3), (9, 4), (0, 1 << 32), ((1 << 32) - 1, 1 << 32), ((1 << 32) - 70000, 1 << 32),
               (123456, 654321), (2 << 16, (2 << 16) + 4097), ((1 << 32) + 1, (1 << 32) + 5),
               # range_end past 2^32: the reference truncates both ends to 32 bits, the closed range is then empty
               (10, (1 << 32) + 5), (100000, (1 << 32) + 5), (1 << 32, (1 << 32) + 7)]


@pytest.mark.parametrize("seed", [17, 18])
def test_oracle_flip_vs_reference(O, R, seed):
    blobs = synth_blobs(R, seed, 40, key_space=6, max_keys=6)
    rng = np.random.default_rng(seed)
    for b in blobs:
        ranges = FLIP_RANGES + [tuple(sorted(rng.integers(0, 7 << 16, 2).tolist())) for _ in range(6)]
        for (s, e) in ranges:
            exp = R.flip_bytes(b, s, e)
            assert O.flip_bytes(b, s, e) == exp, (s, e)
            assert R.flip_bytes(b, s, e, inplace=True) == exp, ("inplace twin differs", s, e)


@pytest.mark.parametrize("seed", [64, 65])
def test_oracle_r64_vs_reference(O, R, seed):
    """64-bit bitmaps through their portable format: the oracle's per-bucket restatement vs
    roaring64_bitmap_{and,or,xor,andnot} (roaring64.c:1332-1895)."""
    from helpers import synth_blobs64
    blobs = synth_blobs64(R, seed, 30)
    rng = np.random.default_rng(seed)
    for _ in range(120):
        i, j = rng.integers(0, len(blobs), 2)
        for op in OPS:
            assert O.r64_op_bytes(op, blobs[i], blobs[j]) == R.r64_op_bytes(op, blobs[i], blobs[j]), (op, i, j)
