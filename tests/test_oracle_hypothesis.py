"""CPU test (-m "not gpu"): property-based pinning of the oracle against the unmodified reference.
hypothesis draws, per bitmap, a handful of (key, container profile, seed, run_optimize) and the
order of the operands; every operation of the path (pairwise + in-place twins, lazy folds with both
bitsetconversion values, or_many / xor_many / or_many_heap, flip, and_cardinality) must give the
reference's bytes.  Shrinking turns a mismatch into a minimal container pair."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import synth

container = st.tuples(st.integers(0, 5), st.sampled_from(synth.PROFILES), st.integers(0, 2 ** 16))
bitmap = st.tuples(st.lists(container, min_size=0, max_size=4, unique_by=lambda c: c[0]), st.booleans())


def build(R, spec):
    conts, ro = spec
    parts = []
    for key, prof, seed in sorted(conts):
        v = np.sort(np.asarray(synth.container_values(np.random.default_rng(seed), prof), dtype=np.uint32))
        parts.append((np.uint32(key) << np.uint32(16)) | v)
    vals = np.concatenate(parts).astype(np.uint32) if parts else np.zeros(0, np.uint32)
    r = R.from_values(vals, run_optimize=ro)
    b = R.serialize(r)
    R.free(r)
    return b


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(specs=st.lists(bitmap, min_size=2, max_size=4), lo=st.integers(0, 7 << 16), span=st.integers(0, 3 << 16))
def test_oracle_equals_reference_on_drawn_bitmaps(O, R, specs, lo, span):
    blobs = [build(R, s) for s in specs]
    a, b = blobs[0], blobs[1]
    for op in ("and", "or", "xor", "andnot"):
        assert O.op_bytes(op, a, b) == R.op_bytes(op, a, b), op
        assert O.op_bytes(op + "_inplace", a, b) == R.op_inplace_bytes(op, a, b), op + "_inplace"
    ra, rb_ = R.deserialize(a), R.deserialize(b)
    assert O.and_cardinality(a, b) == int(R.L.roaring_bitmap_and_cardinality(ra, rb_))
    R.free(ra), R.free(rb_)
    for conv in (False, True):
        assert O.lazy_fold_bytes("or", conv, blobs) == R.lazy_fold_bytes("or", conv, blobs), ("lazy or", conv)
    assert O.lazy_fold_bytes("xor", False, blobs) == R.lazy_fold_bytes("xor", False, blobs), "lazy xor"
    for name in ("or_many", "xor_many"):
        assert O.many_bytes(name, blobs) == R.many_bytes(name, blobs), name
    assert O.or_many_heap_bytes(blobs) == R.many_bytes("or_many_heap", blobs), "or_many_heap"
    assert O.flip_bytes(a, lo, lo + span) == R.flip_bytes(a, lo, lo + span), ("flip", lo, span)
