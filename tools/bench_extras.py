#!/usr/bin/env python
"""The other BASELINE.json configs (bench.py covers configs[1]); one JSON line per workload.

  python tools/bench_extras.py card      # configs[3]: 10^4 bitset-heavy pairs, and_cardinality / jaccard
  python tools/bench_extras.py ormany    # configs[2]: or_many over 200 Zipfian bitmaps, density sweep
  torchrun --nproc-per-node N tools/bench_extras.py sharded   # configs[4]: key-sharded 1000-bitmap OR + NCCL all-reduce

Every line carries the device time measured with CUDA events inside the library on its stream,
algorithmic bytes, fraction of the measured HBM peak, and the unmodified
reference timed on the host cores on the same serialized inputs (checksums compared).
"""
import argparse
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import croaring_b200 as rb  # noqa: E402
from croaring_b200 import sharding as sh  # noqa: E402
from croaring_b200.workloads import bitset_heavy_blobs, zipf_bitmap_blob  # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _gen_bitset(args):
    n, seed = args
    return bitset_heavy_blobs(n, seed=seed)


def _gen_zipf(args):
    seed, n_values, density = args
    return zipf_bitmap_blob(np.random.default_rng(seed), n_values, density)


def compute_ms():
    return float(rb.api.lib().rb200_last_compute_ms())


def bench_card(a):
    """configs[3]: cardinality-only sweep, 10^4 pairs, bitset-heavy (density 0.5, universe 2^20)."""
    npairs = a.pairs
    with Pool(min(32, host_threads())) as p:
        parts = p.map(_gen_bitset, [(500, 1000 + i) for i in range(2 * npairs // 500)])
    blobs = [b for part in parts for b in part]
    t0 = time.perf_counter()
    S = rb.DeviceSet.from_serialized(blobs)
    t_up = time.perf_counter() - t0
    ia = np.arange(0, 2 * npairs, 2, dtype=np.uint32)
    ib = ia + 1
    algo = npairs * 16 * 16384
    ms, tot = [], []
    for it in range(a.warmup + a.steps):
        c = S.and_cardinality(S, ia, ib)
        if it >= a.warmup:
            ms.append(compute_ms())
            tot.append(rb.last_device_ms())
    cards = S.cardinalities()
    jacc = c / (cards[ia] + cards[ib] - c)
    peak, src = peak_gbs()
    k_ms = float(np.median(ms))
    out = {"workload": "cardinality_bitset_heavy (configs[3])", "pairs": npairs,
           "metric": "set-ops/sec (and_cardinality)", "value": npairs / (np.median(tot) * 1e-3),
           "unit": "set-ops/s", "device_ms_per_sweep": float(np.median(tot)), "kernel": "k_card_items",
           "kernel_ms": k_ms, "algorithmic_bytes": algo,
           "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": algo / (k_ms * 1e-3) / 1e9 / peak, "peak_source": src},
           "upload_s": t_up, "sum_and_card": int(c.sum()), "mean_jaccard": float(jacc.mean())}
    if not a.no_cpu:
        from oracle.refbench import RefBench
        rbn = RefBench()
        n_s = min(npairs, 1000)
        h = rbn.load(blobs[:2 * n_s])
        T = host_threads()
        dt1, s1 = rbn.pairs(h, "and_cardinality", ia[:n_s], ib[:n_s], 1)
        dtT, sT = rbn.pairs(h, "and_cardinality", ia[:n_s], ib[:n_s], T)
        assert s1 == sT == int(c[:n_s].sum()), "parity: and_cardinality checksum differs from the reference"
        rbn.unload(h)
        out["cpu_baseline"] = {"kind": "reference", "isa": rbn.isa(), "sample": f"first {n_s} pairs",
                               "value_1thread": n_s / dt1, "gbs_1thread": n_s * 16 * 16384 / dt1 / 1e9,
                               "cores": T, "value": n_s / dtT, "gbs": n_s * 16 * 16384 / dtT / 1e9,
                               "unit": "set-ops/s", "parity": "checksum equal"}
    print(json.dumps(out), flush=True)


def bench_ormany(a):
    """configs[2]: roaring_bitmap_or_many over 200 Zipfian bitmaps, density sweep."""
    peak, src = peak_gbs()
    for d in a.densities:
        with Pool(min(32, host_threads())) as p:
            blobs = p.map(_gen_zipf, [(7000 + b, a.values, d) for b in range(a.bitmaps)])
        S = rb.DeviceSet.from_serialized(blobs)
        ms, tot = [], []
        for it in range(a.warmup + a.steps):
            r = S.or_many()
            if it >= a.warmup:
                ms.append(compute_ms())
                tot.append(rb.last_device_ms())
            card = int(r.cardinalities()[0])
            out_bytes = r.payload_bytes if hasattr(r, "payload_bytes") else 0
            if it < a.warmup + a.steps - 1:
                r.free()
        res = r.download(0)
        res_blob = res.serialize()
        algo = S.payload_bytes + len(res_blob)
        k_ms = float(np.median(ms))
        out = {"workload": "or_many_zipf (configs[2])", "bitmaps": a.bitmaps, "values_per_bitmap": a.values,
               "density": d, "containers": S.container_count, "input_bytes": S.payload_bytes,
               "metric": "set-ops/sec (or_many calls)", "value": 1.0 / (np.median(tot) * 1e-3), "unit": "set-ops/s",
               "input_bitmaps_per_s": a.bitmaps / (np.median(tot) * 1e-3),
               "device_ms": float(np.median(tot)), "kernel": "k_or_many", "kernel_ms": k_ms,
               "algorithmic_bytes": algo, "result_card": card,
               "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": algo / (k_ms * 1e-3) / 1e9 / peak, "peak_source": src}}
        if not a.no_cpu:
            from oracle.refbench import RefBench
            from oracle.refbind import ref
            rbn = RefBench()
            h = rbn.load(blobs)
            dt, c = rbn.or_many(h, np.arange(a.bitmaps, dtype=np.uint32), reps=3)
            rbn.unload(h)
            exp = ref().many_bytes("or_many", blobs)
            out["cpu_baseline"] = {"kind": "reference", "cores": 1, "isa": rbn.isa(), "ms": dt * 1e3,
                                   "value": 1.0 / dt, "unit": "set-ops/s",
                                   "parity": "bytes identical" if exp == res_blob else "MISMATCH"}
            assert c == card
            assert exp == res_blob, "parity: or_many result differs from the reference"
        print(json.dumps(out), flush=True)
        S.free()


def bench_xormany(a):
    """roaring_bitmap_xor_many on the three real-data sets (200 bitmaps each) vs the reference."""
    from oracle.refbind import ref
    import time as _t
    R = ref()
    for ds in ("census1881", "weather_sept_85", "wikileaks-noquotes"):
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        ms = []
        for it in range(a.warmup + a.steps):
            r = S.xor_many()
            if it >= a.warmup:
                ms.append(rb.last_device_ms())
            got = r.serialize_all()[0] if it == a.warmup + a.steps - 1 else None
            r.free()
        rs = [R.deserialize(b) for b in blobs]
        best = 1e9
        for _ in range(5):
            t0 = _t.perf_counter()
            x = R.many("xor_many", rs)
            best = min(best, _t.perf_counter() - t0)
            exp = R.serialize(x)
            R.free(x)
        for x in rs:
            R.free(x)
        print(json.dumps({"workload": "xor_many_realdata", "dataset": ds, "bitmaps": len(blobs),
                          "device_ms": float(np.median(ms)), "value": 1.0 / (np.median(ms) * 1e-3),
                          "unit": "set-ops/s", "input_bytes": S.payload_bytes,
                          "cpu_baseline": {"kind": "reference", "cores": 1, "ms": best * 1e3,
                                           "parity": "bytes identical" if exp == got else "MISMATCH"}}), flush=True)
        assert exp == got


def bench_heap(a):
    """roaring_bitmap_or_many_heap on the three real-data sets (200 bitmaps each) vs the reference.
    Sequential by construction (199 dependent single-pair lazy unions): a parity row, not a
    throughput row — the wall time per call is reported next to the reference's."""
    from oracle.refbind import ref
    import time as _t
    R = ref()
    for ds in ("census1881", "weather_sept_85", "wikileaks-noquotes"):
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        ms = []
        got = None
        for it in range(a.warmup + a.steps):
            rb.synchronize()
            t0 = _t.perf_counter()
            r = S.or_many_heap()
            rb.synchronize()
            if it >= a.warmup:
                ms.append((_t.perf_counter() - t0) * 1e3)
            got = r.serialize_all()[0]
            r.free()
        rs = [R.deserialize(b) for b in blobs]
        best = 1e9
        for _ in range(5):
            t0 = _t.perf_counter()
            x = R.many("or_many_heap", rs)
            best = min(best, _t.perf_counter() - t0)
            exp = R.serialize(x)
            R.free(x)
        for x in rs:
            R.free(x)
        print(json.dumps({"workload": "or_many_heap_realdata", "dataset": ds, "bitmaps": len(blobs),
                          "wall_ms": float(np.median(ms)), "steps_per_call": len(blobs) - 1,
                          "cpu_baseline": {"kind": "reference", "cores": 1, "ms": best * 1e3,
                                           "parity": "bytes identical" if exp == got else "MISMATCH"}}), flush=True)
        assert exp == got


def bench_lazy(a):
    """Batched lazy unions: all 19 900 pairs of a real-data set through roaring_bitmap_lazy_or
    (bitsetconversion off) + roaring_bitmap_repair_after_lazy, device-resident, vs the reference
    doing the same two calls per pair on one host thread (sampled)."""
    from oracle.refbind import ref
    import time as _t
    R = ref()
    for ds in ("census1881", "weather_sept_85", "wikileaks-noquotes"):
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        i, j = np.triu_indices(len(blobs), 1)
        ia, ib = i.astype(np.uint32), j.astype(np.uint32)
        ms = []
        for it in range(a.warmup + a.steps):
            rb.synchronize()
            t0 = _t.perf_counter()
            lz = S.batch("or", S, ia, ib, lazy=True)
            rep = lz.repair_after_lazy()
            rb.synchronize()
            if it >= a.warmup:
                ms.append((_t.perf_counter() - t0) * 1e3)
            cards = rep.cardinalities() if it == a.warmup + a.steps - 1 else None
            lz.free()
            if cards is None:
                rep.free()
        sample = np.arange(0, len(ia), 20)
        rs = [R.deserialize(b) for b in blobs]
        t0 = _t.perf_counter()
        ok = True
        for k in sample:
            x = R.L.roaring_bitmap_lazy_or(rs[ia[k]], rs[ib[k]], False)
            R.L.roaring_bitmap_repair_after_lazy(x)
            ok = ok and int(R.card(x)) == int(cards[k])
            R.free(x)
        cpu_ms = (_t.perf_counter() - t0) * 1e3
        got = rep.serialize_all()
        for k in sample[::10]:
            ok = ok and got[k] == R.lazy_fold_bytes("or", False, [blobs[ia[k]], blobs[ib[k]]])
        rep.free()
        for x in rs:
            R.free(x)
        med = float(np.median(ms))
        print(json.dumps({"workload": "lazy_or+repair all pairs", "dataset": ds, "pairs": int(len(ia)),
                          "wall_ms": med, "value": len(ia) / (med * 1e-3), "unit": "set-ops/s",
                          "cpu_baseline": {"kind": "reference", "cores": 1, "sample_pairs": int(len(sample)),
                                           "value": len(sample) / (cpu_ms * 1e-3),
                                           "parity": "cardinalities + sampled bytes identical" if ok else "MISMATCH"}}),
              flush=True)
        assert ok


def bench_deser(a):
    """Device-side portable deserialization: host blobs -> resident set (H2D of the raw bytes +
    header walk + payload move), vs roaring_bitmap_portable_deserialize_safe on one host thread."""
    from oracle.refbind import ref
    import time as _t
    R = ref()
    for ds in ("census1881", "weather_sept_85", "wikileaks-noquotes"):
        blobs = rb.load_realdata(ds) * 8            # 1600 bitmaps per call
        nbytes = sum(map(len, blobs))
        ms = []
        for it in range(a.warmup + a.steps):
            rb.synchronize()
            t0 = _t.perf_counter()
            S = rb.DeviceSet.from_serialized(blobs)
            rb.synchronize()
            if it >= a.warmup:
                ms.append((_t.perf_counter() - t0) * 1e3)
            if it == a.warmup + a.steps - 1:
                assert S.serialize_all() == blobs
            S.free()
        t0 = _t.perf_counter()
        rs = [R.deserialize(b) for b in blobs]
        cpu_ms = (_t.perf_counter() - t0) * 1e3
        for x in rs:
            R.free(x)
        med = float(np.median(ms))
        print(json.dumps({"workload": "portable deserialize -> resident set", "dataset": ds,
                          "bitmaps": len(blobs), "bytes": nbytes, "wall_ms": med,
                          "GBps": nbytes / (med * 1e-3) / 1e9,
                          "cpu_baseline": {"kind": "reference", "cores": 1, "ms": cpu_ms,
                                           "GBps": nbytes / (cpu_ms * 1e-3) / 1e9,
                                           "parity": "round trip bytes identical"}}), flush=True)


def bench_sharded(a):
    """configs[4]: 10^8-universe, many-bitmap OR sharded by high-16 key range across ranks, one NCCL
    all-reduce of the per-key cardinalities.  Strong scaling: total work fixed."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rb.init(local)
    rb.set_stream(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(12345)
    dens = np.exp(rng.uniform(np.log(0.001), np.log(0.3), size=a.bitmaps))
    with Pool(max(1, min(32, host_threads() // max(world, 1)))) as p:
        blobs = p.map(_gen_zipf, [(9000 + b, int(1e8 * dens[b]), float(dens[b])) for b in range(a.bitmaps)])
    idx = [sh.BlobIndex(b) for b in blobs]
    ranges = sh.plan_key_ranges(sh.key_byte_histogram(idx), world)
    lo, hi = ranges[rank]
    mine = [sh.slice_blob_by_keys(ix, lo, hi) for ix in idx]
    S = rb.DeviceSet.from_serialized(mine)
    dev = torch.device("cuda", local)
    times, kms = [], []
    cards = None
    for it in range(a.warmup + a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cpk = np.zeros(65536, dtype=np.uint32)
        r = S.or_many(key_lo=lo, key_hi=hi, card_per_key=cpk)
        cards = sh.allreduce_cardinalities(cpk, dist if world > 1 else None, dev)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= a.warmup:
            times.append(float(t.item()))
            kms.append(compute_ms())
        if it < a.warmup + a.steps - 1:
            r.free()
    shard_blob = r.download(0).serialize()
    in_bytes = torch.tensor([S.payload_bytes], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(in_bytes)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(shard_blob, parts, dst=0)
    else:
        parts = [shard_blob]
    if rank == 0:
        full = sh.concat_blobs(parts)
        peak, src = peak_gbs()
        ms = float(np.median(times))
        total_in = int(in_bytes.item())
        out = {"workload": "or_many_sharded (configs[4])", "n_gpus": world, "bitmaps": a.bitmaps, "universe": 10 ** 8,
               "scaling": "strong", "key_ranges": ranges, "metric": "set-ops/sec (or_many calls)",
               "value": 1.0 / (ms * 1e-3), "unit": "set-ops/s", "input_bitmaps_per_s": a.bitmaps / (ms * 1e-3),
               "ms_per_call_max_over_ranks": ms, "kernel_ms_rank0": float(np.median(kms)),
               "input_bytes_all_ranks": total_in, "result_card": int(cards.sum()), "result_bytes": len(full),
               "collective": "1 x all_reduce(sum) of int64[65536] per-key cardinalities (NCCL)" if world > 1 else "none (1 rank)",
               "roofline": {"bound": "hbm", "achieved": total_in / (ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                            "frac": total_in / (ms * 1e-3) / 1e9 / (peak * world), "peak_source": src}}
        if not a.no_cpu:
            from oracle.refbench import RefBench
            from oracle.refbind import ref
            rbn = RefBench()
            h = rbn.load(blobs)
            dt, c = rbn.or_many(h, np.arange(a.bitmaps, dtype=np.uint32), reps=1)
            rbn.unload(h)
            exp = ref().many_bytes("or_many", blobs)
            out["cpu_baseline"] = {"kind": "reference", "cores": 1, "ms": dt * 1e3, "value": 1.0 / dt,
                                   "unit": "set-ops/s", "parity": "bytes identical" if exp == full else "MISMATCH"}
            assert c == int(cards.sum()) and exp == full, "parity: sharded or_many differs from the reference"
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["card", "ormany", "sharded", "xormany", "heap", "lazy", "deser"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--bitmaps", type=int, default=200)
    ap.add_argument("--values", type=int, default=10 ** 6)
    ap.add_argument("--densities", type=float, nargs="+", default=[0.3, 0.1, 0.03, 0.01, 0.003])
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.what != "sharded":
        rb.init(0)
    {"card": bench_card, "ormany": bench_ormany, "sharded": bench_sharded, "xormany": bench_xormany,
     "heap": bench_heap, "lazy": bench_lazy, "deser": bench_deser}[a.what](a)
