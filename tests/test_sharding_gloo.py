"""CPU tests (-m "not gpu") of the N>1 path: key-range planning, host slicing / concatenation of
portable bitmaps, and the world_size-2 gloo run of the sharded many-way OR (the per-rank
reduction is played by the oracle here; on GPU ranks it is DeviceSet.or_many)."""
import os
import socket
import sys

import numpy as np
import pytest

import croaring_b200.datasets as dsm
from croaring_b200 import sharding as sh
from helpers import synth_blobs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blob_index_roundtrip(R):
    blobs = dsm.load_realdata("wikileaks-noquotes")[:30] + dsm.load_realdata("weather_sept_85")[:5] \
        + synth_blobs(R, 3, 30)
    for b in blobs:
        ix = sh.BlobIndex(b)
        mv = memoryview(b)
        pay = [mv[int(s):int(s + z)] for s, z in zip(ix.starts, ix.sizes)]
        assert sh.build_blob(ix.keys, ix.cards, ix.isrun, pay) == b
        # slicing into 3 key ranges and concatenating gives the bitmap back
        parts = [sh.slice_blob_by_keys(b, lo, hi) for lo, hi in ((0, 2), (3, 40), (41, 65535))]
        assert sh.concat_blobs(parts) == b
        for p in parts:                       # every slice is a valid bitmap for the reference
            r = R.deserialize(p)
            assert R.validate(r)[0]
            R.free(r)


def test_plan_key_ranges_balanced_and_covering():
    rng = np.random.default_rng(0)
    hist = np.zeros(65536, dtype=np.int64)
    hist[:1526] = (rng.pareto(1.0, 1526) * 1e5).astype(np.int64) + 1     # Zipf-ish head
    for world in (1, 2, 4, 8):
        rs = sh.plan_key_ranges(hist, world)
        assert rs[0][0] == 0 and rs[-1][1] == 65535
        for (a, b), (c, d) in zip(rs, rs[1:]):
            assert b + 1 == c and a <= b and c <= d
        loads = [hist[a:b + 1].sum() for a, b in rs]
        assert max(loads) <= hist.sum() / world + hist.max() + 1


def _worker(rank, world, port, seed, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle.oraclebind import oracle
    from oracle.refbind import ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O, R = oracle(), ref()
    blobs = dsm.load_realdata("census1881")[:40] + synth_blobs(R, seed, 40, key_space=24, max_keys=12)

    def engine(shard_blobs, lo, hi):
        out = O.many_bytes("or_many", shard_blobs)
        ix = sh.BlobIndex(out)
        cpk = np.zeros(65536, dtype=np.uint32)
        cpk[ix.keys.astype(np.int64)] = ix.cards
        assert len(ix.keys) == 0 or (ix.keys.min() >= lo and ix.keys.max() <= hi)
        return out, cpk

    full, cards, (lo, hi) = sh.or_many_sharded(blobs, rank, world, engine, dist=dist)
    exp = O.many_bytes("or_many", blobs)
    total = O.cardinality(exp)
    ok = int(cards.sum()) == total                      # every rank knows the total after the all-reduce
    if rank == 0:
        ok = ok and full == exp and R.many_bytes("or_many", blobs) == exp
    q.put((rank, ok, lo, hi))
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
