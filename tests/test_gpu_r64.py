"""GPU parity of the 64-bit path (roaring64.c:1332-1895 through the 64-bit portable format)."""
import numpy as np
import pytest

from helpers import OPS, synth_blobs64

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [81, 82])
def test_r64_batch_ops(rb, R, O, seed):
    blobs = synth_blobs64(R, seed, 40)
    rng = np.random.default_rng(seed)
    ia = rng.integers(0, len(blobs), 300).astype(np.uint32)
    ib = rng.integers(0, len(blobs), 300).astype(np.uint32)
    for op in OPS:
        got = rb.r64_batch_op(op, blobs, blobs, ia, ib)
        for k in range(len(ia)):
            exp = R.r64_op_bytes(op, blobs[ia[k]], blobs[ib[k]])
            assert got[k] == exp, (op, k)
        assert O.r64_op_bytes(op, blobs[ia[0]], blobs[ib[0]]) == got[0]
    cards = rb.r64_and_cardinality(blobs, blobs, ia, ib)
    for k in range(0, len(ia), 3):
        assert int(cards[k]) == R.r64_and_cardinality(blobs[ia[k]], blobs[ib[k]])


def test_r64_edge_cases(rb, R):
    empty = R.r64_from_values(np.zeros(0, np.uint64))
    one = R.r64_from_values(np.array([5, (7 << 32) + 9, (1 << 63) + 1], dtype=np.uint64))
    big = R.r64_from_values((np.arange(200000, dtype=np.uint64) * 3) + (np.uint64(7) << np.uint64(32)))
    blobs = [empty, one, big]
    ia = np.array([0, 0, 1, 1, 2, 2, 1, 2], dtype=np.uint32)
    ib = np.array([0, 1, 0, 1, 1, 2, 2, 0], dtype=np.uint32)
    for op in OPS:
        got = rb.r64_batch_op(op, blobs, blobs, ia, ib)
        for k in range(len(ia)):
            assert got[k] == R.r64_op_bytes(op, blobs[ia[k]], blobs[ib[k]]), (op, k)
    with pytest.raises(rb.RB200Error):
        rb.r64_batch_op("or", [one[:-3]], blobs, ia[:1], ib[:1])
