"""GPU parity of roaring_bitmap_flip / flip_inplace (roaring.h:986-1010; negation cells of
mixed_negation.c) and of unusual INPUT representations (shared containers, frozen views)."""
import numpy as np
import pytest

from helpers import check_result_bitmap, synth_blobs
from test_oracle_pinning import FLIP_RANGES

pytestmark = pytest.mark.gpu

PROFILES = ["full", "nearfull", "halves", "dense", "bitset", "array", "tiny", "longruns", "shortruns", "ends"]


@pytest.mark.parametrize("seed", [51, 52])
def test_batch_flip_vs_reference(rb, R, O, seed):
    blobs = synth_blobs(R, seed, 60, key_space=6, max_keys=6, profiles=PROFILES if seed == 52 else None)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(seed)
    ranges = FLIP_RANGES + [tuple(sorted(rng.integers(0, 7 << 16, 2).tolist())) for _ in range(8)]
    for (s, e) in ranges:
        got = S.flip(s, e).serialize_all()
        for k, b in enumerate(blobs):
            exp = R.flip_bytes(b, s, e)
            assert got[k] == exp, (s, e, k)
        assert O.flip_bytes(blobs[0], s, e) == got[0]
    idx = np.array([3, 3, 17, 0], dtype=np.uint32)
    sub = S.flip(70000, 200000, idx).download_all()
    for k, i in enumerate(idx):
        check_result_bitmap(R, sub[k], R.flip_bytes(blobs[i], 70000, 200000), f"idx {i}")


def test_flip_dropin_and_inplace(rb, R):
    blobs = synth_blobs(R, 53, 12, key_space=4, max_keys=5, profiles=PROFILES)
    for b in blobs:
        for (s, e) in [(0, 1 << 32), (65536, 65537), (1000, 300000), (7, 7)]:
            exp = R.flip_bytes(b, s, e)
            x = rb.Bitmap.deserialize(b)
            check_result_bitmap(R, x.flip(s, e), exp, f"flip {s} {e}")
            x.flip_inplace(s, e)
            check_result_bitmap(R, x, exp, f"flip_inplace {s} {e}")


def test_shared_and_frozen_inputs(rb, R):
    """Inputs the reference may hand us: COW bitmaps whose containers are SHARED
    (containers.h:71-75) and FROZEN views (roaring.c:3401) — read-only memory, never mutated."""
    blobs = synth_blobs(R, 54, 16, key_space=5, max_keys=6)
    for i in range(0, 16, 2):
        a, b = R.deserialize(blobs[i]), R.deserialize(blobs[i + 1])
        R.L.roaring_bitmap_set_copy_on_write(a, True)
        a2 = R.L.roaring_bitmap_copy(a)           # containers of a / a2 are now shared wrappers
        fb = R.frozen_bytes(blobs[i + 1])
        import ctypes as C
        raw = C.create_string_buffer(len(fb) + 64)
        base = (C.addressof(raw) + 31) & ~31
        C.memmove(base, fb, len(fb))
        view = R.L.roaring_bitmap_frozen_view(base, len(fb))
        assert view
        L = rb.lib()
        for op in ("and", "or", "xor", "andnot"):
            exp = R.op_bytes(op, blobs[i], blobs[i + 1])
            out = rb.Bitmap(getattr(L, f"roaring_bitmap_{op}")(a2, view))
            assert out.serialize() == exp, (i, op)
            ok, why = R.validate(out.ptr)
            assert ok, why
        assert int(L.roaring_bitmap_and_cardinality(a2, view)) == int(R.L.roaring_bitmap_and_cardinality(a, b))
        assert R.serialize(a2) == blobs[i] and R.serialize(view) == blobs[i + 1]   # inputs untouched
        for x in (a, a2, b, view):
            R.free(x)
