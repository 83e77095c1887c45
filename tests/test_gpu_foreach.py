"""GPU test of the visitor download (rb200_download_foreach)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_foreach_sum_cardinality(rb, golden):
    blobs = rb.load_realdata("weather_sept_85")
    S = rb.DeviceSet.from_serialized(blobs)
    ia = np.arange(199, dtype=np.uint32)
    for op in ("and", "or", "xor"):
        r = S.batch(op, S, ia, ia + 1)
        assert r.foreach_sum_cardinality() == golden["weather_sept_85"]["run_optimized"][op]["sum_card"]
    e = S.batch("and", S, np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    assert e.foreach_sum_cardinality() == 0
