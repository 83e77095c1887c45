"""Parity pin of the TIMED bench workload: all 19 900 unordered pairs of each
bench dataset, and / or / xor / andnot, against golden values produced by the unmodified reference
(tests/golden/make_allpairs_golden.py): sum of cardinalities, total portable size and the sha256
over the concatenated portable serialisations (device-side serialization of every result)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "allpairs_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("ds", ["census1881", "weather_sept_85", "wikileaks-noquotes"])
def test_allpairs_bytes_match_reference_golden(rb, gold, ds):
    blobs = rb.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    ia, ib = np.triu_indices(len(blobs), 1)
    ia, ib = ia.astype(np.uint32), ib.astype(np.uint32)
    assert len(ia) == gold[ds]["pairs"]
    for op in ("and", "or", "xor", "andnot"):
        r = S.batch(op, S, ia, ib)
        assert int(r.cardinalities().sum()) == gold[ds][op]["sum_card"], (ds, op)
        buf, off, ln, release = r.serialize_all(copy=False)
        h = hashlib.sha256()
        total = 0
        base = buf.value
        for k in range(len(ia)):
            n = int(ln[k])
            h.update(C.string_at(base + int(off[k]), n))
            total += n
        release()
        r.free()
        assert total == gold[ds][op]["sum_portable_bytes"], (ds, op)
        assert h.hexdigest() == gold[ds][op]["sha256"], (ds, op)
    c = S.and_cardinality(S, ia, ib)
    assert int(c.sum()) == gold[ds]["and_cardinality"]


def test_allpairs_sharded_checksum(rb, gold):
    """The strong-scaling split of bench.py (rank r owns pairs r, r+N, ...): the per-rank device
    checksums add up to the reference's checksum for any N."""
    tot = 0
    for ds in ("census1881", "wikileaks-noquotes"):
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        ia, ib = np.triu_indices(len(blobs), 1)
        ia, ib = ia.astype(np.uint32), ib.astype(np.uint32)
        for op in ("and", "or", "xor"):
            for world in (3,):
                s = 0
                for rank in range(world):
                    r = S.batch(op, S, ia[rank::world].copy(), ib[rank::world].copy())
                    s += int(r.cardinalities().sum())
                    r.free()
                assert s == gold[ds][op]["sum_card"]
            tot += s
    assert tot == sum(gold[ds][op]["sum_card"] for ds in ("census1881", "wikileaks-noquotes") for op in ("and", "or", "xor"))

