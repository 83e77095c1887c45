"""Key-range sharding of many-way unions across GPUs (BASELINE.json configs[4]).

Python mirror of the C entry points in include/roaring_b200.h ("Multi-GPU"): the planning /
slicing / concatenation of portable bitmaps is host C (rb200_plan_key_ranges,
rb200_blob_slice_keys, rb200_blobs_concat), the per-rank reduction and the ONE collective
(ncclAllReduce of uint32[K] per-key cardinalities, on the device) are rb200_or_many_sharded.

  rank g:  ranges, span = plan_key_ranges(blobs, world)          # identical on every rank
           S = DeviceSet.from_serialized(blobs, *ranges[g])      # only the containers of range g
           part, cards, total = S.or_many_sharded(comm, *ranges[g], span)
  result = concatenation of the parts in rank order (concat_blobs)

`or_many_sharded_host` is the same protocol with the per-rank reduction and the collective
passed in — the CPU tests play it over gloo with the oracle as the reduction.
"""
import numpy as np

from . import api


def plan_key_ranges(blobs, world):
    """([(lo, hi)] * world contiguous key ranges covering 0..65535 balanced by container bytes,
    (first, last) live key)."""
    return api.plan_key_ranges(blobs, world)


def slice_blob_by_keys(blob: bytes, key_lo: int, key_hi: int) -> bytes:
    return api.blob_slice_keys(blob, key_lo, key_hi)


def concat_blobs(blobs) -> bytes:
    return api.blobs_concat(blobs)


def or_many_sharded(blobs, comm, gather=None):
    """GPU form: this rank's part of roaring_bitmap_or_many(blobs) as portable bytes, the per-key
    cardinalities of the span (every rank holds the same array after the all-reduce), the total
    cardinality and (ranges, span).  gather(part_bytes) -> list of all parts on rank 0 (else
    None) turns it into the full result on rank 0."""
    ranges, span = plan_key_ranges(blobs, comm.size)
    lo, hi = ranges[comm.rank]
    S = api.DeviceSet.from_serialized(blobs, lo, hi)
    R, cards, total = S.or_many_sharded(comm, lo, hi, span)
    part = R.serialize_all()[0]
    R.free()
    S.free()
    full = None
    if gather is not None:
        parts = gather(part)
        if parts is not None:
            full = concat_blobs(parts)
    return (full if gather is not None else part), cards, total, (ranges, span)


def or_many_sharded_host(blobs, rank, world, reduce_range, allreduce_u32):
    """The protocol with injected pieces (CPU tests): reduce_range(shard_blobs, lo, hi) -> portable
    bytes of this rank's union; allreduce_u32(np.uint32[K]) -> summed array."""
    ranges, span = plan_key_ranges(blobs, world)
    lo, hi = ranges[rank]
    mine = [slice_blob_by_keys(b, lo, hi) for b in blobs]
    part = reduce_range(mine, lo, hi)
    K = max(0, span[1] - span[0] + 1)
    cards = np.zeros(K, dtype=np.uint32)
    keys, kcards = blob_key_cards(part)
    if len(keys):
        assert keys.min() >= lo and keys.max() <= hi
        cards[keys.astype(np.int64) - span[0]] = kcards
    cards = allreduce_u32(cards) if world > 1 else cards
    return part, cards, (ranges, span)


def blob_key_cards(blob: bytes):
    """(keys u16, cardinalities u32) from the descriptive header of a portable bitmap."""
    cookie = int(np.frombuffer(blob[:4], dtype="<u4")[0])
    if (cookie & 0xFFFF) == 12347:
        n = (cookie >> 16) + 1
        pos = 4 + (n + 7) // 8
    else:
        n = int(np.frombuffer(blob[4:8], dtype="<u4")[0])
        pos = 8
    kc = np.frombuffer(blob[pos:pos + 4 * n], dtype="<u2").reshape(n, 2)
    return kc[:, 0].copy(), kc[:, 1].astype(np.uint32) + 1
