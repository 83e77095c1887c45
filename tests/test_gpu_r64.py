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


_INMEM = r'''
import ctypes as C, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
# the host application's CRoaring: loaded GLOBAL, as in any process that owns roaring64_bitmap_t objects
ref_so = os.path.join(sys.argv[1], "oracle", "_ref", "libroaring_ref.so")
G = C.CDLL(ref_so, mode=os.RTLD_GLOBAL)
from oracle.refbind import ref
import croaring_b200 as rb
from helpers import synth_blobs64, OPS
R = ref()
rb.init(0)
blobs = synth_blobs64(R, 83, 24)
bms = [R.r64_deserialize(b) for b in blobs]
rng = np.random.default_rng(3)
ia = rng.integers(0, len(bms), 60); ib = rng.integers(0, len(bms), 60)
for op in OPS:
    outs = rb.api.r64_batch_op_inmemory(op, [bms[i] for i in ia], [bms[j] for j in ib])
    for k, o in enumerate(outs):
        why = C.c_char_p()
        assert R.L.roaring64_bitmap_internal_validate(o, C.byref(why)), why.value
        assert R.r64_serialize(o) == R.r64_op_bytes(op, blobs[ia[k]], blobs[ib[k]]), (op, k)
        R.L.roaring64_bitmap_free(o)
    # the drop-in symbol on the host library's objects
    o = getattr(rb.lib(), f"roaring64_bitmap_{op}")(bms[ia[0]], bms[ib[0]])
    assert o and R.r64_serialize(o) == R.r64_op_bytes(op, blobs[ia[0]], blobs[ib[0]])
    R.L.roaring64_bitmap_free(o)
L = rb.lib()
for k in range(0, 60, 7):
    a, b = bms[ia[k]], bms[ib[k]]
    inter = R.r64_and_cardinality(blobs[ia[k]], blobs[ib[k]])
    ca, cb = R.L.roaring64_bitmap_get_cardinality(a), R.L.roaring64_bitmap_get_cardinality(b)
    assert L.roaring64_bitmap_and_cardinality(a, b) == inter
    assert L.roaring64_bitmap_or_cardinality(a, b) == ca + cb - inter
    assert L.roaring64_bitmap_xor_cardinality(a, b) == ca + cb - 2 * inter
    assert L.roaring64_bitmap_andnot_cardinality(a, b) == ca - inter
    assert bool(L.roaring64_bitmap_intersect(a, b)) == (inter > 0)
print("r64 in-memory ok")
'''


def test_r64_in_memory_binding(tmp_path):
    """roaring64_bitmap_t objects of the host application's CRoaring in, roaring64_bitmap_t out
    (rb200_r64_batch_op + the roaring64_bitmap_* drop-in symbols), in a process where the
    reference is loaded globally — the deployment the binding is made for."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", _INMEM, root], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "r64 in-memory ok" in p.stdout, (p.stdout[-500:], p.stderr[-1500:])
