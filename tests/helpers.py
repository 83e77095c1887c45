"""Shared helpers for the parity tests."""
import hashlib

import numpy as np

import synth

DATASETS = ["census1881", "weather_sept_85", "wikileaks-noquotes",
            "census1881_srt", "wikileaks-noquotes_srt", "uscensus2000"]
OPS = ["and", "or", "xor", "andnot"]


def synth_blobs(R, seed, n, key_space=10, max_keys=9, profiles=None):
    """n seeded synthetic bitmaps, serialized by the reference (mixed run-optimized or not)."""
    rng = np.random.default_rng(seed)
    blobs = []
    for i in range(n):
        vals = synth.random_bitmap(rng, n_keys=int(rng.integers(0, max_keys)), key_space=key_space,
                                   profiles=profiles)
        r = R.from_values(vals, run_optimize=bool(i % 3))
        blobs.append(R.serialize(r))
        R.free(r)
    return blobs


def no_run_twins(R, blobs):
    out = []
    for b in blobs:
        r = R.deserialize(b)
        R.L.roaring_bitmap_remove_run_compression(r)
        out.append(R.serialize(r))
        R.free(r)
    return out


def sha_concat(blobs):
    h = hashlib.sha256()
    for b in blobs:
        h.update(b)
    return h.hexdigest()


def check_result_bitmap(R, bm, expect_bytes, what=""):
    """bm: croaring_b200.Bitmap produced by the CUDA path. Byte-exact + valid for the reference."""
    got = bm.serialize()
    assert got == expect_bytes, f"{what}: portable bytes differ (ours {len(got)} B, ref {len(expect_bytes)} B)"
    ok, why = R.validate(bm.ptr)           # the reference's own validator on OUR object
    assert ok, f"{what}: reference validate failed: {why}"
    assert R.serialize(bm.ptr) == expect_bytes, f"{what}: reference serializer disagrees on our object"
