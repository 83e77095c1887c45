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
