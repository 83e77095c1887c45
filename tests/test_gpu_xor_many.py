"""GPU parity of roaring_bitmap_xor_many: the sequential lazy-xor fold and
its type transitions vs the unmodified reference and the oracle, byte-exact."""
import numpy as np
import pytest

from helpers import DATASETS, check_result_bitmap, no_run_twins, sha_concat, synth_blobs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ds", DATASETS)
def test_realdata_xor_many(rb, R, golden, ds):
    blobs = rb.load_realdata(ds)
    S = rb.DeviceSet.from_serialized(blobs)
    res = S.xor_many()
    out = res.download(0)
    g = golden[ds]["run_optimized"]["xor_many"]
    assert out.cardinality() == g["card"] == int(res.cardinalities()[0])
    assert sha_concat([out.serialize()]) == g["sha256"]
    check_result_bitmap(R, out, R.many_bytes("xor_many", blobs), f"{ds} xor_many")
    nr = no_run_twins(R, blobs)
    out2 = rb.DeviceSet.from_serialized(nr).xor_many().download(0)
    assert sha_concat([out2.serialize()]) == golden[ds]["no_runs"]["xor_many"]["sha256"]


@pytest.mark.parametrize("seed", [41, 42, 43])
def test_synthetic_xor_many(rb, R, O, seed):
    blobs = synth_blobs(R, seed, 60, key_space=6, max_keys=7)
    S = rb.DeviceSet.from_serialized(blobs)
    rng = np.random.default_rng(seed)
    for trial in range(50):
        n = int(rng.integers(0, 10))
        idx = rng.integers(0, len(blobs), n).astype(np.uint32)
        if trial % 7 == 0 and n >= 2:
            idx[1] = idx[0]                      # x ^ x: keys emptied, then re-inserted by later inputs
        out = S.xor_many(idx).download(0)
        sub = [blobs[i] for i in idx]
        exp = R.many_bytes("xor_many", sub)
        assert O.many_bytes("xor_many", sub) == exp
        check_result_bitmap(R, out, exp, f"seed {seed} trial {trial} idx {idx.tolist()}")


def test_dropin_xor_many(rb, R):
    blobs = rb.load_realdata("wikileaks-noquotes")[:9]
    bms = [rb.Bitmap.deserialize(b) for b in blobs]
    check_result_bitmap(R, rb.xor_many(bms), R.many_bytes("xor_many", blobs), "dropin xor_many")
