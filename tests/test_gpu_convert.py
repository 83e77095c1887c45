"""GPU tests of the batch producers / consumers next to the path:
run_optimize / remove_run_compression and to_uint32_array vs the unmodified reference."""
import numpy as np
import pytest

from helpers import synth_blobs

pytestmark = pytest.mark.gpu


def _mk(R, seed, n):
    rng = np.random.default_rng(seed)
    import synth
    out = []
    for i in range(n):
        r = R.from_values(synth.random_bitmap(rng, n_keys=int(rng.integers(0, 9)), key_space=10), run_optimize=False)
        out.append(r)
    return out


def test_run_optimize_and_remove(rb, R):
    raw = _mk(R, 3, 80)
    plain = [R.serialize(r) for r in raw]
    S = rb.DeviceSet.from_serialized(plain)
    opt = S.run_optimize().serialize_all()
    for r in raw:
        R.L.roaring_bitmap_run_optimize(r)
    exp_opt = [R.serialize(r) for r in raw]
    assert opt == exp_opt
    # and back: remove_run_compression of the optimized set
    S2 = rb.DeviceSet.from_serialized(exp_opt)
    back = S2.run_optimize(remove_runs=True).serialize_all()
    for r in raw:
        R.L.roaring_bitmap_remove_run_compression(r)
    assert back == [R.serialize(r) for r in raw]
    # idempotent, and usable as op input
    again = rb.DeviceSet.from_serialized(opt).run_optimize().serialize_all()
    assert again == opt
    for r in raw:
        R.free(r)


def test_to_uint32_arrays(rb, R):
    blobs = rb.load_realdata("weather_sept_85")[:30] + rb.load_realdata("wikileaks-noquotes")[:30] + synth_blobs(R, 9, 40)
    S = rb.DeviceSet.from_serialized(blobs)
    got = S.to_uint32_arrays()
    for k, b in enumerate(blobs):
        r = R.deserialize(b)
        exp = R.to_array(r)
        R.free(r)
        assert got[k].dtype == np.uint32 and np.array_equal(got[k], exp), k
    # results of an op decode too
    ia = np.arange(len(blobs) - 1, dtype=np.uint32)
    res = S.batch("xor", S, ia, ia + 1)
    vals = res.to_uint32_arrays()
    for k in (0, 17, 60):
        e = R.deserialize(R.op_bytes("xor", blobs[k], blobs[k + 1]))
        assert np.array_equal(vals[k], R.to_array(e))
        R.free(e)
