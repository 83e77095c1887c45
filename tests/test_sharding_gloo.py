"""CPU tests (-m "not gpu") of the N>1 path: the host C entry points for key-range planning,
slicing and concatenation of portable bitmaps (rb200_plan_key_ranges / rb200_blob_slice_keys /
rb200_blobs_concat — no CUDA involved), and the world_size-2 gloo run of the sharded many-way OR
protocol (the per-rank reduction is played by the oracle here and the all-reduce of the uint32[K]
per-key cardinalities by gloo; on GPU ranks they are rb200_or_many_sharded + NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

import croaring_b200.datasets as dsm
from croaring_b200 import sharding as sh
from croaring_b200 import workloads as wl
from helpers import synth_blobs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_concat_roundtrip(R):
    blobs = dsm.load_realdata("wikileaks-noquotes")[:30] + dsm.load_realdata("weather_sept_85")[:5] \
        + synth_blobs(R, 3, 30)
    for b in blobs:
        # slicing into 3 key ranges and concatenating gives the bitmap back
        parts = [sh.slice_blob_by_keys(b, lo, hi) for lo, hi in ((0, 2), (3, 40), (41, 65535))]
        assert sh.concat_blobs(parts) == b
        assert sh.slice_blob_by_keys(b, 0, 65535) == b
        for p, (lo, hi) in zip(parts, ((0, 2), (3, 40), (41, 65535))):
            r = R.deserialize(p)                 # every slice is a valid bitmap for the reference
            assert R.validate(r)[0]
            vals = R.to_array(r)
            assert len(vals) == 0 or (int(vals[0]) >> 16 >= lo and int(vals[-1]) >> 16 <= hi)
            R.free(r)
    with pytest.raises(Exception):
        sh.concat_blobs([blobs[0], blobs[0]])    # overlapping key ranges are refused


def test_plan_key_ranges_balanced_and_covering():
    A = wl.zipf_arena(24, 60 * 65536, None, density_draw=True)       # Zipf: bytes concentrate in low keys
    blobs = A.blobs()
    A.free()
    hist = np.zeros(65536, dtype=np.int64)
    for b in blobs:
        keys, _ = sh.blob_key_cards(b)
        for k in keys:
            hist[k] += len(sh.slice_blob_by_keys(b, int(k), int(k))) + 512  # header + payload of that key + the per-container weight
    for world in (1, 2, 3, 4, 8):
        rs, span = sh.plan_key_ranges(blobs, world)
        assert span == (0, 59)
        assert rs[0][0] == 0 and rs[-1][1] == 65535
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b + 1 == c and a <= b and c <= d
        loads = [hist[a:b + 1].sum() for a, b in rs]
        assert max(loads) <= hist.sum() / world + hist.max() + 1, (world, loads)


def _worker(rank, world, port, seed, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.oraclebind import oracle
    from oracle.refbind import ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O, R = oracle(), ref()
    blobs = dsm.load_realdata("census1881")[:40] + synth_blobs(R, seed, 40, key_space=24, max_keys=12)

    def allreduce(a):
        t = torch.from_numpy(a.astype(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy().astype(np.uint32)

    part, cards, (ranges, span) = sh.or_many_sharded_host(
        blobs, rank, world, lambda mine, lo, hi: O.many_bytes("or_many", mine), allreduce)
    exp = O.many_bytes("or_many", blobs)
    ok = int(cards.astype(np.int64).sum()) == O.cardinality(exp)   # every rank knows the total after the all-reduce
    ekeys, ecards = sh.blob_key_cards(exp)
    ok = ok and len(cards) == span[1] - span[0] + 1 and np.array_equal(cards[ekeys.astype(np.int64) - span[0]], ecards)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    if rank == 0:
        full = sh.concat_blobs(parts)
        ok = ok and full == exp and R.many_bytes("or_many", blobs) == exp
    q.put((rank, bool(ok), ranges[rank][0], ranges[rank][1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("seed", [7, 8])
def test_or_many_sharded_world2_gloo(seed):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    ranges = sorted((lo, hi) for _, _, lo, hi in res)
    assert ranges[0][0] == 0 and ranges[1][1] == 65535 and ranges[0][1] + 1 == ranges[1][0]


def _blob(R, values, run_optimize=False):
    r = R.from_values(np.asarray(values, dtype=np.uint32), run_optimize=run_optimize)
    b = R.serialize(r)
    R.free(r)
    return b


def test_blob_algebra_edge_cases(R):
    """Edges of the host blob algebra: empty bitmaps, run cookies with fewer than four containers
    (no offset header, roaring_array.c:500-506), slices that keep nothing, more ranks than live
    keys, malformed inputs."""
    empty = _blob(R, [])
    assert len(empty) == 8
    assert sh.slice_blob_by_keys(empty, 0, 65535) == empty
    assert sh.concat_blobs([empty, empty, empty]) == empty
    assert sh.concat_blobs([]) == empty
    # three run containers: run cookie WITHOUT offsets; slicing to two keeps the short header form,
    # concatenating two such bitmaps (4 containers) must grow the offset header
    runs3 = _blob(R, list(range(0, 5000)) + list(range(1 << 16, (1 << 16) + 7000)) + list(range(5 << 16, (5 << 16) + 4500)), True)
    cookie = int(np.frombuffer(runs3[:4], dtype="<u4")[0])
    assert (cookie & 0xFFFF) == 12347 and (cookie >> 16) + 1 == 3
    two = sh.slice_blob_by_keys(runs3, 0, 1)
    rt = R.deserialize(two)
    assert R.validate(rt)[0] and R.card(rt) == 12000
    assert R.serialize(rt) == two                       # canonical bytes (what the reference would write)
    R.free(rt)
    other = _blob(R, list(range(9 << 16, (9 << 16) + 6000)) + [(11 << 16) + 5, (11 << 16) + 9], True)
    cat = sh.concat_blobs([runs3, other])               # 3 + 2 containers, mixed run / array
    rc = R.deserialize(cat)
    assert R.validate(rc)[0] and R.card(rc) == 5000 + 7000 + 4500 + 6000 + 2
    assert R.serialize(rc) == cat
    R.free(rc)
    # a slice that keeps nothing is the canonical empty bitmap
    assert sh.slice_blob_by_keys(runs3, 2, 4) == empty
    assert sh.slice_blob_by_keys(runs3, 6, 65535) == empty
    # more ranks than live keys: ranges still cover 0..65535, are disjoint and non-empty as key ranges
    rs, span = sh.plan_key_ranges([runs3, other], 8)
    assert span == (0, 11) and rs[0][0] == 0 and rs[-1][1] == 65535
    for (a, b), (c, d) in zip(rs, rs[1:]):
        assert a <= b and b + 1 == c and c <= d
    parts = [sh.concat_blobs([sh.slice_blob_by_keys(x, lo, hi) for x in (runs3, other)]) for lo, hi in rs]
    assert sh.concat_blobs(parts) == cat
    # no container at all: an empty span
    rs, span = sh.plan_key_ranges([empty, empty], 2)
    assert span[0] > span[1] and rs[0][0] == 0 and rs[-1][1] == 65535
    # malformed inputs are refused, not read out of bounds
    for bad in (b"", b"\x00\x01", runs3[:-1], runs3[:10], b"\xff" * 16):
        with pytest.raises(Exception):
            sh.slice_blob_by_keys(bad, 0, 65535)
        with pytest.raises(Exception):
            sh.plan_key_ranges([runs3, bad], 2)
        with pytest.raises(Exception):
            sh.concat_blobs([bad])
