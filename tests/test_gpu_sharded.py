"""GPU tests of the key-sharded many-way OR: the ranks of a G-GPU job are
played one after the other on cuda:0 — each "rank" holds only its key range of every input,
runs rb200_or_many_keyrange, and the per-key cardinalities are summed as the all-reduce would."""
import numpy as np
import pytest

from croaring_b200 import sharding as sh
from croaring_b200.workloads import zipf_blobs
from helpers import synth_blobs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_or_many_matches_reference(rb, R, world):
    blobs = rb.load_realdata("weather_sept_85")[:50] + synth_blobs(R, 31, 50, key_space=30, max_keys=14) \
        + zipf_blobs(12, 300000, 0.1, seed=3)
    ranges, span = sh.plan_key_ranges(blobs, world)
    one = rb.Comm.create(0, 1)          # a 1-rank communicator: rb200_or_many_sharded issues no collective
    total = np.zeros(span[1] - span[0] + 1, dtype=np.int64)
    shards = []
    for rank in range(world):
        lo, hi = ranges[rank]
        S = rb.DeviceSet.from_serialized(blobs, lo, hi)          # host slicing in C + streamed upload
        part, cards, tot = S.or_many_sharded(one, lo, hi, span)
        assert tot == int(cards.astype(np.int64).sum())
        total += cards                                           # what the NCCL all-reduce(sum) computes
        shards.append(part.serialize_all()[0])
    one.destroy()
    exp = R.many_bytes("or_many", blobs)
    assert sh.concat_blobs(shards) == exp
    e = R.deserialize(exp)
    assert int(total.sum()) == R.card(e)
    R.free(e)
    ekeys, ecards = sh.blob_key_cards(exp)
    assert np.array_equal(total[ekeys.astype(np.int64) - span[0]], ecards)


def test_keyrange_on_unsliced_inputs(rb, R):
    """rb200_or_many_keyrange restricted to a key window on full inputs == window of the full result."""
    blobs = synth_blobs(R, 77, 40, key_space=16, max_keys=10)
    S = rb.DeviceSet.from_serialized(blobs)
    exp = R.many_bytes("or_many", blobs)
    for lo, hi in ((0, 3), (4, 9), (10, 65535), (5, 5)):
        cpk = np.zeros(65536, dtype=np.uint32)
        got = S.or_many(key_lo=lo, key_hi=hi, card_per_key=cpk).download(0).serialize()
        assert got == sh.slice_blob_by_keys(exp, lo, hi)
        assert int(cpk.sum()) == int(sh.blob_key_cards(got)[1].sum())


def test_zipf_or_many_saturation(rb, R, O):
    """configs[2] shape (scaled): Zipfian inputs whose low keys saturate -> full-run state machine."""
    for d in (0.3, 0.03):
        blobs = zipf_blobs(24, 400000, d, seed=11)
        S = rb.DeviceSet.from_serialized(blobs)
        got = S.or_many().download(0).serialize()
        exp = R.many_bytes("or_many", blobs)
        assert got == exp
        assert O.many_bytes("or_many", blobs) == exp
