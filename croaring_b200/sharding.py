"""Key-range sharding of many-way unions across GPUs (BASELINE.json configs[4]).

The high-16 key space is cut into G contiguous ranges balanced by input bytes; rank g holds, for
every input bitmap, only the containers whose key falls in range g (host-side slicing of the
portable bytes — no device-to-device traffic), reduces its keys independently on its GPU, and
the result is the concatenation of the per-rank results in rank order.  The only collective is
one all-reduce(sum) of the per-key result cardinalities (uint32[65536], zeros outside the own
range) so that every rank knows the total cardinality / per-key sizes.

Pure host logic (numpy + torch.distributed); the device work is DeviceSet.or_many.  The parsing
helpers restate the portable format (/root/reference/src/roaring_array.c:469-531).
"""
import numpy as np

SERIAL_COOKIE_NO_RUN = 12346
SERIAL_COOKIE = 12347
NO_OFFSET_THRESHOLD = 4


class BlobIndex:
    """Header-level view of one portable bitmap: keys, cardinalities, run flags, payload spans."""

    __slots__ = ("blob", "keys", "cards", "isrun", "starts", "sizes")

    def __init__(self, blob: bytes):
        self.blob = blob
        mv = memoryview(blob)
        cookie = int(np.frombuffer(mv[:4], dtype="<u4")[0])
        pos = 4
        if (cookie & 0xFFFF) == SERIAL_COOKIE:
            n = (cookie >> 16) + 1
            nb = (n + 7) // 8
            flags = np.frombuffer(mv[pos:pos + nb], dtype=np.uint8)
            isrun = np.unpackbits(flags, bitorder="little")[:n].astype(bool)
            pos += nb
            hasrun = True
        elif cookie == SERIAL_COOKIE_NO_RUN:
            n = int(np.frombuffer(mv[4:8], dtype="<u4")[0])
            pos = 8
            isrun = np.zeros(n, dtype=bool)
            hasrun = False
        else:
            raise ValueError("not a portable roaring bitmap")
        kc = np.frombuffer(mv[pos:pos + 4 * n], dtype="<u2").reshape(n, 2)
        pos += 4 * n
        if (not hasrun) or n >= NO_OFFSET_THRESHOLD:
            pos += 4 * n
        self.keys = kc[:, 0].astype(np.uint16)
        self.cards = kc[:, 1].astype(np.int64) + 1
        self.isrun = isrun
        starts = np.zeros(n, dtype=np.int64)
        sizes = np.zeros(n, dtype=np.int64)
        p = pos
        for i in range(n):
            starts[i] = p
            if isrun[i]:
                nr = int(np.frombuffer(mv[p:p + 2], dtype="<u2")[0])
                sizes[i] = 2 + 4 * nr
            elif self.cards[i] > 4096:
                sizes[i] = 8192
            else:
                sizes[i] = 2 * self.cards[i]
            p += int(sizes[i])
        self.starts, self.sizes = starts, sizes


def build_blob(keys, cards, isrun, payloads):
    """Portable serialization from container-level pieces (payloads: list of bytes-like)."""
    n = len(keys)
    hasrun = bool(np.any(isrun)) if n else False
    parts = []
    if hasrun:
        parts.append(np.array([SERIAL_COOKIE | ((n - 1) << 16)], dtype="<u4").tobytes())
        parts.append(np.packbits(np.asarray(isrun, dtype=np.uint8), bitorder="little").tobytes())
        hdr = 4 + (n + 7) // 8 + (4 * n if n < NO_OFFSET_THRESHOLD else 8 * n)
    else:
        parts.append(np.array([SERIAL_COOKIE_NO_RUN, n], dtype="<u4").tobytes())
        hdr = 8 + 8 * n
    kc = np.zeros((n, 2), dtype="<u2")
    if n:
        kc[:, 0] = np.asarray(keys, dtype=np.uint16)
        kc[:, 1] = (np.asarray(cards, dtype=np.int64) - 1).astype(np.uint16)
    parts.append(kc.tobytes())
    sizes = np.array([len(p) for p in payloads], dtype=np.int64)
    if (not hasrun) or n >= NO_OFFSET_THRESHOLD:
        offs = hdr + np.concatenate([[0], np.cumsum(sizes)[:-1]]) if n else np.zeros(0)
        parts.append(np.asarray(offs, dtype="<u4").tobytes())
    parts.extend(bytes(p) for p in payloads)
    return b"".join(parts)


def slice_blob_by_keys(blob: bytes, key_lo: int, key_hi: int) -> bytes:
    """The bitmap restricted to containers with key in [key_lo, key_hi]."""
    ix = blob if isinstance(blob, BlobIndex) else BlobIndex(blob)
    sel = np.flatnonzero((ix.keys >= key_lo) & (ix.keys <= key_hi))
    mv = memoryview(ix.blob)
    pay = [mv[int(ix.starts[i]):int(ix.starts[i] + ix.sizes[i])] for i in sel]
    return build_blob(ix.keys[sel], ix.cards[sel], ix.isrun[sel], pay)


def concat_blobs(blobs):
    """Concatenate bitmaps living on disjoint, increasing key ranges (per-rank results)."""
    keys, cards, isrun, pay = [], [], [], []
    last = -1
    for b in blobs:
        ix = BlobIndex(b)
        if len(ix.keys):
            if int(ix.keys[0]) <= last:
                raise ValueError("shards are not on increasing disjoint key ranges")
            last = int(ix.keys[-1])
        mv = memoryview(b)
        keys.append(ix.keys)
        cards.append(ix.cards)
        isrun.append(ix.isrun)
        pay.extend(mv[int(s):int(s + z)] for s, z in zip(ix.starts, ix.sizes))
    if not keys:
        return build_blob([], [], [], [])
    return build_blob(np.concatenate(keys), np.concatenate(cards), np.concatenate(isrun), pay)


def key_byte_histogram(indexes):
    """Input bytes per high-16 key summed over all bitmaps (BlobIndex list)."""
    h = np.zeros(65536, dtype=np.int64)
    for ix in indexes:
        np.add.at(h, ix.keys.astype(np.int64), ix.sizes)
    return h


def plan_key_ranges(byte_hist, world):
    """`world` contiguous key ranges [lo, hi] covering 0..65535, balanced by input bytes
    (prefix sum of the key histogram; Zipfian data concentrates bytes in low keys)."""
    c = np.cumsum(byte_hist.astype(np.float64))
    total = c[-1] if c[-1] > 0 else 1.0
    bounds = [0]
    for g in range(1, world):
        k = int(np.searchsorted(c, total * g / world, side="left")) + 1
        k = min(max(k, bounds[-1] + 1), 65536 - (world - g))
        bounds.append(k)
    bounds.append(65536)
    return [(bounds[g], bounds[g + 1] - 1) for g in range(world)]


def allreduce_cardinalities(card_per_key, dist=None, device=None):
    """Sum the per-key cardinality arrays (uint32[65536]) over all ranks; returns int64 numpy.
    One collective per many-way op, only when world > 1 (NCCL on GPU ranks, gloo in CPU tests)."""
    out = card_per_key.astype(np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    import torch
    t = torch.from_numpy(out)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def or_many_sharded(blobs, rank, world, engine_or_many, dist=None, device=None, gather=True):
    """Many-way OR of serialized bitmaps, key-sharded over `world` ranks.

    engine_or_many(shard_blobs, key_lo, key_hi) -> (result_blob, card_per_key uint32[65536]) is
    the per-rank reduction (the CUDA engine on GPU ranks).  Returns (full result blob on rank 0
    when gather else own shard blob, per-key cardinalities known to every rank, my (lo, hi))."""
    idx = [BlobIndex(b) for b in blobs]
    ranges = plan_key_ranges(key_byte_histogram(idx), world)
    lo, hi = ranges[rank]
    mine = [slice_blob_by_keys(ix, lo, hi) for ix in idx]
    shard, cpk = engine_or_many(mine, lo, hi)
    cards = allreduce_cardinalities(cpk, dist, device)
    if not gather or world == 1 or dist is None:
        return shard, cards, (lo, hi)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(shard, parts, dst=0)
    full = concat_blobs(parts) if rank == 0 else None
    return full, cards, (lo, hi)
