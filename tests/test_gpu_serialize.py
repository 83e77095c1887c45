"""GPU tests of the device-side portable serialization: bytes produced on
the device equal the reference's roaring_bitmap_portable_serialize of the same results."""
import numpy as np
import pytest

from helpers import OPS, synth_blobs

pytestmark = pytest.mark.gpu


def test_serialize_inputs_roundtrip(rb, R):
    blobs = rb.load_realdata("wikileaks-noquotes")[:50] + rb.load_realdata("weather_sept_85")[:20] \
        + synth_blobs(R, 61, 60)
    empty = R.serialize(R.L.roaring_bitmap_create_with_capacity(0))
    blobs.append(empty)
    S = rb.DeviceSet.from_serialized(blobs)
    assert S.serialize_all() == blobs


@pytest.mark.parametrize("ds", ["census1881", "wikileaks-noquotes"])
def test_serialize_batch_results(rb, R, ds):
    blobs = rb.load_realdata(ds)[:80]
    S = rb.DeviceSet.from_serialized(blobs)
    i, j = np.triu_indices(len(blobs), 1)
    for op in OPS:
        res = S.batch(op, S, i.astype(np.uint32), j.astype(np.uint32))
        got = res.serialize_all()
        for k in range(0, len(got), 97):
            assert got[k] == R.op_bytes(op, blobs[i[k]], blobs[j[k]]), (ds, op, k)
        assert got == [b.serialize() for b in res.download_all()]
