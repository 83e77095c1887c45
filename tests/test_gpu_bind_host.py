"""GPU test: host-bound sets elide pass-through payloads from the download; results identical."""
import numpy as np
import pytest

from helpers import OPS, synth_blobs

pytestmark = pytest.mark.gpu


def test_bind_host_results_identical_and_smaller(rb, R):
    blobs = rb.load_realdata("census1881")[:60] + synth_blobs(R, 91, 40)
    host = [rb.Bitmap.deserialize(b) for b in blobs]
    i, j = np.triu_indices(len(blobs), 1)
    ia, ib = i.astype(np.uint32), j.astype(np.uint32)
    S0 = rb.DeviceSet.upload(host)
    S1 = rb.DeviceSet.upload(host).bind_host()
    for op in OPS:
        r0, r1 = S0.batch(op, S0, ia, ib), S1.batch(op, S1, ia, ib)
        a = [x.serialize() for x in r0.download_all()]
        full_bytes = rb.api.lib().rb200_last_download_bytes()
        b = [x.serialize() for x in r1.download_all()]
        elided_bytes = rb.api.lib().rb200_last_download_bytes()
        assert a == b
        for k in range(0, len(a), 211):
            assert a[k] == R.op_bytes(op, blobs[i[k]], blobs[j[k]])
        if op != "and":
            assert elided_bytes < full_bytes          # pass-through payloads stayed home
        assert r1.foreach_sum_cardinality() == int(r0.cardinalities().sum())
        # chained op on a bound result still works (payloads exist on the device)
        k = np.arange(len(ia), dtype=np.uint32)
        c0 = r0.batch("xor", S0, k, ib)
        c1 = r1.batch("xor", S1, k, ib)
        assert [x.serialize() for x in c0.download_all()][::97] == [x.serialize() for x in c1.download_all()][::97]


def test_only_one_parent_bound(rb, R):
    """A result whose LEFT parent is host-bound and whose right parent is not (and vice versa):
    pass-through containers of the unbound side travel over PCIe, the bound side's are rebuilt."""
    import numpy as np
    from helpers import synth_blobs
    blobs = synth_blobs(R, 71, 30, key_space=8, max_keys=8)
    host = [rb.Bitmap.deserialize(b) for b in blobs]
    A = rb.DeviceSet.upload(host).bind_host()
    B = rb.DeviceSet.from_serialized(blobs)
    ia = np.arange(29, dtype=np.uint32)
    for left, right in ((A, B), (B, A)):
        for op in ("or", "xor", "andnot"):
            res = left.batch(op, right, ia, ia + 1)
            exp = [R.op_bytes(op, blobs[k], blobs[k + 1]) for k in range(29)]
            assert [b.serialize() for b in res.download_all()] == exp
            assert res.download(5).serialize() == exp[5]
