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


def test_foreach_many_pipelined(rb, R, golden):
    """Several result sets through one pipelined stream: built-in visitor checksum, then a Python
    visitor that keeps every bitmap and checks index -> content against the reference."""
    import threading
    sets, exp_sum, exp_bytes = [], 0, []
    keep = []
    for ds in ("wikileaks-noquotes", "census1881"):
        blobs = rb.load_realdata(ds)[:60]
        host = [rb.Bitmap.deserialize(b) for b in blobs]
        keep.append(host)
        S = rb.DeviceSet.upload(host).bind_host()
        ia = np.arange(59, dtype=np.uint32)
        for op in ("and", "or", "xor", "andnot"):
            sets.append(S.batch(op, S, ia, ia + 1))
            for k in range(59):
                b = R.op_bytes(op, blobs[k], blobs[k + 1])
                exp_bytes.append(b)
    exp_sum = sum(int(s.cardinalities().sum()) for s in sets)
    assert rb.foreach_many(sets) == exp_sum
    got, lock = {}, threading.Lock()

    def visit(i, p):
        bm = rb.Bitmap(p)                      # we keep it (return 1): freed by the wrapper later
        with lock:
            got[i] = bm
        return 1

    rb.foreach_many(sets, visit)
    assert sorted(got) == list(range(len(exp_bytes)))
    for i, b in enumerate(exp_bytes):
        assert got[i].serialize() == b, i
    assert rb.foreach_many([]) == 0


def test_foreach_async_overlaps_and_matches(rb, R):
    """Background downloader: results queued while the caller keeps launching ops; checksums and
    contents equal the synchronous path; sets are freed right after being queued."""
    import ctypes as C
    import threading
    blobs = rb.load_realdata("weather_sept_85")[:50]
    host = [rb.Bitmap.deserialize(b) for b in blobs]
    i, j = np.triu_indices(len(blobs), 1)
    ia, ib = i.astype(np.uint32), j.astype(np.uint32)
    exp = {}
    S = rb.DeviceSet.upload(host).bind_host()
    for op in ("and", "or", "xor", "andnot"):
        exp[op] = int(S.batch(op, S, ia, ib).cardinalities().sum())
    for rep in range(3):
        acc = rb.CardinalitySum()
        for op in ("and", "or", "xor", "andnot"):
            r = S.batch(op, S, ia, ib)
            r.foreach_async(acc)
            r.free()                                  # the job owns a packed copy
        e = S.batch("and", S, np.zeros(0, np.uint32), np.zeros(0, np.uint32))
        e.foreach_async(acc)                          # empty set: a job with no chunks
        rb.download_wait()
        assert acc.value == sum(exp.values())
    # custom visitor: keep the bitmaps, compare bytes
    got, lock = {}, threading.Lock()

    def visit(idx, p, _ctx):
        with lock:
            got[idx] = rb.Bitmap(p)
        return 1

    cb = rb.api.VISIT_FN(visit)
    r = S.batch("xor", S, ia[:300], ib[:300])
    r.foreach_async(None, cb)
    rb.download_wait()
    assert sorted(got) == list(range(300))
    for k in range(0, 300, 7):
        assert got[k].serialize() == R.op_bytes("xor", blobs[ia[k]], blobs[ib[k]])
    rb.download_wait()                                # nothing queued: returns at once
