"""GPU parity of the in-place twins: roaring_bitmap_
{and,or,xor,andnot}_inplace through the drop-in symbols, and the in-place type rules in batches."""
import numpy as np
import pytest

from helpers import OPS, synth_blobs

pytestmark = pytest.mark.gpu


def test_inplace_dropins(rb, R, O):
    blobs = rb.load_realdata("weather_sept_85")[:8] + synth_blobs(R, 71, 30, key_space=6, max_keys=7,
                                                                  profiles=["full", "nearfull", "halves", "dense",
                                                                            "bitset", "array", "longruns", "tiny"])
    rng = np.random.default_rng(5)
    for _ in range(60):
        i, j = rng.integers(0, len(blobs), 2)
        for op in OPS:
            x = rb.Bitmap.deserialize(blobs[i])
            y = rb.Bitmap.deserialize(blobs[j])
            x.inplace(op, y)
            exp = R.op_inplace_bytes(op, blobs[i], blobs[j])
            assert x.serialize() == exp, (op, i, j)
            assert O.op_bytes(op + "_inplace", blobs[i], blobs[j]) == exp
            ok, why = R.validate(x.ptr)
            assert ok, why
            assert y.serialize() == blobs[j]          # right operand untouched


def test_inplace_or_full_rules(rb, R):
    """even|odd -> RUN in place but BITSET functionally; full left container kept as is (T5)."""
    even = np.arange(0, 65536, 2, dtype=np.uint32)
    odd = np.arange(1, 65536, 2, dtype=np.uint32)
    full = np.arange(65536, dtype=np.uint32)
    mk = lambda v, ro: R.serialize(R.from_values(v, run_optimize=ro))
    cases = [(mk(even, False), mk(odd, False)), (mk(full, False), mk(even, False)),
             (mk(full, True), mk(even, False)), (mk(even, False), mk(full, True)),
             (mk(even, False), mk(np.array([1, 3], dtype=np.uint32), False))]
    blobs = [b for c in cases for b in c]
    S = rb.DeviceSet.from_serialized(blobs)
    ia = np.arange(0, len(blobs), 2, dtype=np.uint32)
    res = S.batch("or", S, ia, ia + 1, inplace_rules=True).download_all()
    fun = S.batch("or", S, ia, ia + 1).download_all()
    for k, (a, b) in enumerate(cases):
        assert res[k].serialize() == R.op_inplace_bytes("or", a, b), k
        assert fun[k].serialize() == R.op_bytes("or", a, b), k
    assert res[0].serialize() != fun[0].serialize()       # run vs bitset


def test_inplace_with_shared_left_containers(rb, R):
    """COW bitmaps: after roaring_bitmap_copy of a copy-on-write bitmap every
    container is a SHARED wrapper; the reference's in-place twins then take the FUNCTIONAL cell for
    a shared left container (src/roaring.c:840, 1085-1088, 1235, 1376) — observable in the
    saturating bitset | bitset cell (BITSET, not the full RUN of container_ior)."""
    even = np.arange(0, 65536, 2, dtype=np.uint32)
    odd = np.arange(1, 65536, 2, dtype=np.uint32)
    mk = lambda v, ro: R.serialize(R.from_values(v, run_optimize=ro))
    extra = synth_blobs(R, 91, 12, key_space=5, max_keys=6, profiles=["full", "nearfull", "halves", "dense", "bitset", "array"])
    lefts = [mk(np.concatenate([even, even + (1 << 16)]), False)] + extra[:6]
    rights = [mk(np.concatenate([odd, odd + (1 << 16)]), False)] + extra[6:]
    seen_diff = False
    for a, b in zip(lefts, rights):
        for op in OPS:
            outs = []
            for who in ("ref", "ours"):
                x = R.deserialize(a)
                R.L.roaring_bitmap_set_copy_on_write(x, True)
                y = R.L.roaring_bitmap_copy(x)             # x and y now share every container
                rhs = R.deserialize(b)
                if who == "ref":
                    getattr(R.L, f"roaring_bitmap_{op}_inplace")(y, rhs)
                else:
                    getattr(rb.lib(), f"roaring_bitmap_{op}_inplace")(y, rhs)   # OUR symbol on THEIR object
                ok, why = R.validate(y)
                assert ok, (who, op, why)
                outs.append(R.serialize(y))
                assert R.serialize(x) == a                # the sharing sibling is untouched
                for z in (x, y, rhs):
                    R.free(z)                             # the reference tears down what we swapped in
            assert outs[0] == outs[1], op
            if op == "or" and outs[0] != R.op_inplace_bytes("or", a, b):
                seen_diff = True                          # shared left: functional cell, differs from plain in-place
    assert seen_diff
