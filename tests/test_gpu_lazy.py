"""GPU parity of the public lazy API (roaring.h:932-977) and roaring_bitmap_or_many_heap
(roaring_priority_queue.c:200) against the reference and the oracle."""
import numpy as np
import pytest

import croaring_b200.datasets as dsm
from helpers import check_result_bitmap, synth_blobs

pytestmark = pytest.mark.gpu

PROFILES = ["full", "nearfull", "halves", "dense", "bitset", "array", "tiny", "longruns",
            "shortruns", "ends"]


def _fold_device(rb, S, idx, op, conv):
    """Left fold of lazy ops on the device (every step = one single-pair batch), then repair."""
    one = lambda v: np.array([v], dtype=np.uint32)
    if len(idx) == 1:
        E = rb.DeviceSet.from_serialized([S_EMPTY])
        return S.batch("or", E, one(idx[0]), one(0), lazy=True).repair_after_lazy()
    acc = S.batch(op, S, one(idx[0]), one(idx[1]), lazy=True, bitsetconversion=conv)
    for k in idx[2:]:
        acc = acc.batch(op, S, one(0), one(k), lazy=True, inplace_rules=True, bitsetconversion=conv)
    return acc.repair_after_lazy()


S_EMPTY = None


@pytest.fixture(autouse=True)
def _empty_blob(R):
    global S_EMPTY
    if S_EMPTY is None:
        r = R.from_values(np.zeros(0, np.uint32), False)
        S_EMPTY = R.serialize(r)
        R.free(r)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_lazy_fold_device(rb, R, O, seed):
    blobs = synth_blobs(R, seed, 50, key_space=6, max_keys=7, profiles=PROFILES if seed != 13 else None)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(seed)
    for trial in range(40):
        idx = rng.integers(0, len(blobs), int(rng.integers(1, 8))).tolist()
        sub = [blobs[i] for i in idx]
        for op, conv in (("or", False), ("or", True), ("xor", False)):
            exp = R.lazy_fold_bytes(op, conv, sub)
            assert O.lazy_fold_bytes(op, conv, sub) == exp
            out = _fold_device(rb, S, idx, op, conv).download(0)
            check_result_bitmap(R, out, exp, f"seed {seed} trial {trial} {op} conv={conv} idx {idx}")


def test_lazy_batched_pairs(rb, R):
    """Many pairs per launch under the lazy rules, repaired as one set."""
    blobs = synth_blobs(R, 21, 80, key_space=8, max_keys=8, profiles=PROFILES)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(21)
    ia = rng.integers(0, len(blobs), 500).astype(np.uint32)
    ib = rng.integers(0, len(blobs), 500).astype(np.uint32)
    for op, conv in (("or", False), ("or", True), ("xor", False)):
        lz = S.batch(op, S, ia, ib, lazy=True, bitsetconversion=conv)
        with pytest.raises(rb.RB200Error):
            lz.batch("and", S, ia[:1], ib[:1])          # lazy state: refused until repaired
        got = lz.repair_after_lazy().serialize_all()
        for k in range(len(ia)):
            exp = R.lazy_fold_bytes(op, conv, [blobs[ia[k]], blobs[ib[k]]])
            assert got[k] == exp, (op, conv, k)


def test_lazy_dropin_symbols_mix_with_reference(rb, R):
    """Our lazy results are host bitmaps in the reference's lazy representation: the reference's
    own lazy functions and repair accept them, and ours accept the reference's."""
    blobs = synth_blobs(R, 31, 30, key_space=5, max_keys=6, profiles=PROFILES)
    rng = np.random.default_rng(31)
    for trial in range(25):
        i, j, k = rng.integers(0, len(blobs), 3)
        for conv in (False, True):
            exp = R.lazy_fold_bytes("or", conv, [blobs[i], blobs[j], blobs[k]])
            a, b, c = (rb.Bitmap.deserialize(blobs[t]) for t in (i, j, k))
            # ours all the way
            acc = a.lazy_or(b, conv).lazy_or_inplace(c, conv).repair_after_lazy()
            check_result_bitmap(R, acc, exp, f"ours {trial} conv={conv}")
            # ours lazy_or -> the reference's lazy_or_inplace and repair on OUR object
            mid = a.lazy_or(b, conv)
            R.L.roaring_bitmap_lazy_or_inplace(mid.ptr, c.ptr, conv)
            R.L.roaring_bitmap_repair_after_lazy(mid.ptr)
            assert R.serialize(mid.ptr) == exp
            # the reference's lazy_or -> our lazy_or_inplace and repair on ITS object
            ra, rb_, rc = (R.deserialize(blobs[t]) for t in (i, j, k))
            rmid = R.L.roaring_bitmap_lazy_or(ra, rb_, conv)
            rb.lib().roaring_bitmap_lazy_or_inplace(rmid, rc, conv)
            rb.lib().roaring_bitmap_repair_after_lazy(rmid)
            assert R.serialize(rmid) == exp
            for x in (ra, rb_, rc, rmid):
                R.free(x)
        exp = R.lazy_fold_bytes("xor", False, [blobs[i], blobs[j], blobs[k]])
        a, b, c = (rb.Bitmap.deserialize(blobs[t]) for t in (i, j, k))
        acc = a.lazy_xor(b).lazy_xor_inplace(c).repair_after_lazy()
        check_result_bitmap(R, acc, exp, f"xor {trial}")


@pytest.mark.parametrize("seed", [41, 42])
def test_or_many_heap_synthetic(rb, R, O, seed):
    blobs = synth_blobs(R, seed, 60, key_space=6, max_keys=7, profiles=PROFILES)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(seed)
    for trial in range(30):
        n = int(rng.integers(0, 14))
        idx = rng.integers(0, len(blobs), n).astype(np.uint32)
        sub = [blobs[i] for i in idx]
        exp = R.many_bytes("or_many_heap", sub)
        assert O.or_many_heap_bytes(sub) == exp
        out = S.or_many_heap(idx).download(0)
        check_result_bitmap(R, out, exp, f"heap seed {seed} trial {trial} idx {idx.tolist()}")


@pytest.mark.parametrize("ds", ["census1881", "weather_sept_85", "wikileaks-noquotes"])
def test_or_many_heap_realdata(rb, R, ds):
    blobs = dsm.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    check_result_bitmap(R, S.or_many_heap().download(0), R.many_bytes("or_many_heap", blobs), ds)
    host = [rb.Bitmap.deserialize(b) for b in blobs[:40]]
    check_result_bitmap(R, rb.or_many_heap(host), R.many_bytes("or_many_heap", blobs[:40]), ds + " drop-in")
