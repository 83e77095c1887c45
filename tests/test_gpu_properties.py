"""GPU property tests at BASELINE.json-scale shapes (no reference needed at these sizes):
idempotence, sharding consistency, inclusion-exclusion — size-independent invariants of the path."""
import numpy as np
import pytest

from croaring_b200 import sharding as sh
from croaring_b200.workloads import zipf_blobs

pytestmark = pytest.mark.gpu


def test_or_many_config2_scale_properties(rb):
    """configs[2] shape: 200 Zipfian bitmaps x 10^6 values."""
    blobs = zipf_blobs(200, 10 ** 6, 0.03, seed=21)
    S = rb.DeviceSet.from_serialized(blobs)
    full = S.or_many()
    out = full.serialize_all()[0]
    card = int(full.cardinalities()[0])
    cards = S.cardinalities()
    assert card >= int(cards.max()) and card <= int(cards.sum())
    # idempotence: every input twice, in another order, gives the same VALUE set; and since the
    # fold is order dependent only through container types, the cardinality per key is identical
    idx = np.concatenate([np.arange(200), np.arange(199, -1, -1)]).astype(np.uint32)
    twice = S.or_many(idx)
    assert int(twice.cardinalities()[0]) == card
    # x | or_many == or_many (pairwise op against the many-way result, byte level on values)
    R1 = rb.DeviceSet.from_serialized([out])
    ia = np.arange(200, dtype=np.uint32)
    again = S.batch("or", R1, ia, np.zeros(200, np.uint32))
    assert (again.cardinalities() == card).all()
    inter = S.and_cardinality(R1, ia, np.zeros(200, np.uint32))
    assert (inter == cards).all()                      # every input is a subset of the union
    # key-sharded evaluation concatenates to the same bytes
    ranges, _span = sh.plan_key_ranges(blobs, 4)
    parts, tot = [], np.zeros(65536, dtype=np.int64)
    for lo, hi in ranges:
        cpk = np.zeros(65536, dtype=np.uint32)
        parts.append(S.or_many(key_lo=lo, key_hi=hi, card_per_key=cpk).serialize_all()[0])
        tot += cpk
    assert sh.concat_blobs(parts) == out
    assert int(tot.sum()) == card


def test_xor_many_pairs_cancel(rb):
    """xor_many(x0, x0, x1, x1, ...) is empty; xor_many(x, y) == x ^ y."""
    blobs = zipf_blobs(40, 200000, 0.05, seed=5)
    S = rb.DeviceSet.from_serialized(blobs)
    idx = np.repeat(np.arange(40), 2).astype(np.uint32)
    e = S.xor_many(idx)
    assert int(e.cardinalities()[0]) == 0
    assert S.xor_many(np.array([3, 7], dtype=np.uint32)).serialize_all()[0] == \
        S.batch("xor", S, np.array([3], np.uint32), np.array([7], np.uint32)).serialize_all()[0]
