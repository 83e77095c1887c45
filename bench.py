#!/usr/bin/env python
"""bench.py — set-ops/sec of the Roaring hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (C ABI, sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...   # unmodified CRoaring on host cores

Workload (config.workload): the reference's real-data suite
census1881 + weather_sept_85 + wikileaks-noquotes (200 run-optimized bitmaps each, committed as
portable-serialized fixtures), ops AND / OR / XOR on ALL 19 900 unordered bitmap pairs of each
set, one batched call per (dataset, op): 179 100 set-ops per step.  This is the scaled form of
BASELINE.json configs[1]; the literal 199-successive-pairs sweep (config 2a) is reported beside
it under "successive".

One JSON line on stdout (rank 0).  `value` = whole-job set-ops/s with inputs resident in HBM;
`e2e` = same metric through the C ABI with HOST roaring_bitmap_t in and out (upload + kernels +
download + host materialisation inside the timed region).  See the contract in DESIGN.md §6.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DATASETS = ["census1881", "weather_sept_85", "wikileaks-noquotes"]
OPS = ["and", "or", "xor"]
METRIC = "set-ops/sec (AND/OR/XOR over realdata suite)"


_REAL_STDOUT = None


def capture_stdout():
    """Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are
    sent to stderr for the rest of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def all_pairs(n):
    i, j = np.triu_indices(n, 1)
    return i.astype(np.uint32), j.astype(np.uint32)


def successive_pairs(n):
    return np.arange(n - 1, dtype=np.uint32), np.arange(1, n, dtype=np.uint32)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """Unmodified CRoaring (oracle/_ref) on the host cores, same workload / metric / unit."""
    if rank != 0:
        return
    from oracle.refbench import RefBench
    import croaring_b200.datasets as dsm
    rbn = RefBench()
    T = host_threads()
    sets = {ds: rbn.load(dsm.load_realdata(ds)) for ds in DATASETS}
    pairs = {ds: all_pairs(sets[ds][1]) for ds in DATASETS}
    ops_per_step = sum(len(pairs[ds][0]) for ds in DATASETS) * len(OPS)

    def step():
        t = 0.0
        for ds in DATASETS:
            for op in OPS:
                dt, _ = rbn.pairs(sets[ds], op, pairs[ds][0], pairs[ds][1], T)
                t += dt
        return t

    for _ in range(args.warmup):
        step()
    tot = sum(step() for _ in range(args.steps))
    val = ops_per_step * args.steps / tot
    # the literal configs[1] sweep (199 successive pairs) for context, best of 5, <= 16 threads
    succ_t = 1e9
    for _ in range(5):
        t = 0.0
        for ds in DATASETS:
            ia, ib = successive_pairs(sets[ds][1])
            for op in OPS:
                t += rbn.pairs(sets[ds], op, ia, ib, min(T, 16))[0]
        succ_t = min(succ_t, t)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "set-ops/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16/u64 bitwise", "data": "reference realdata fixtures",
        "config": {"workload": "realdata_allpairs", "datasets": DATASETS, "ops": OPS,
                   "pairs_per_dataset": int(len(pairs[DATASETS[0]][0])),
                   "set_ops_per_step": ops_per_step},
        "cpu_baseline": {"value": val, "unit": "set-ops/s", "cores": T, "kind": "reference",
                         "isa": rbn.isa(),
                         "sample": "full workload per step (all 19 900 pairs x 3 ops x 3 datasets), "
                                   "pairs split statically over all host threads"},
        "e2e": {"value": val, "unit": "set-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "successive": {"workload": "realdata_successive (199 pairs x 3 ops x 3 datasets)",
                       "value": 3 * 3 * 199 / succ_t, "unit": "set-ops/s", "threads": min(T, 16)},
    }
    emit(line)


# ------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    capture_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import croaring_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rb.init(local_rank)
    stream = torch.cuda.current_stream()
    rb.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: portable-serialized fixtures -> device-resident sets (outside the timed region)
    blobs = {ds: rb.load_realdata(ds) for ds in DATASETS}
    sets = {ds: rb.DeviceSet.from_serialized(blobs[ds]) for ds in DATASETS}
    pairs = {ds: all_pairs(len(blobs[ds])) for ds in DATASETS}
    succ = {ds: successive_pairs(len(blobs[ds])) for ds in DATASETS}
    ops_per_step = sum(len(pairs[ds][0]) for ds in DATASETS) * len(OPS)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    stats = {"algo_bytes": 0, "kernel_ms": 0.0, "launches": 0, "checksum": 0}

    def step(pairset, collect=False):
        # the nine batch calls queue back to back on the stream (a call returns once its kernels
        # are enqueued); results are then consumed: counters, per-result cardinalities (D2H), free
        res = []
        for ds in DATASETS:
            ia, ib = pairset[ds]
            for op in OPS:
                res.append(sets[ds].batch(op, sets[ds], ia, ib))
        chk = 0
        for r in res:
            if collect:
                _, cms, ab = r.op_stats()
                stats["algo_bytes"] += ab
                stats["kernel_ms"] += cms
                stats["launches"] += 1
                chk += int(r.cardinalities().sum())
            r.free()
        return chk

    def timed(pairset, steps, warmup, collect):
        for _ in range(warmup):
            step(pairset)
        barrier()
        tot_ms = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(steps):
            flush_buf.zero_()            # flush L2 between timed iterations (not timed)
            torch.cuda.synchronize()
            e0.record(stream)
            chk = step(pairset, collect)
            e1.record(stream)
            torch.cuda.synchronize()
            tot_ms += e0.elapsed_time(e1)
            if collect:
                stats["checksum"] = chk
        barrier()
        t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = rb.kernel_launches()
    tot_ms = timed(pairs, args.steps, args.warmup, True)
    gpu_launches = rb.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * ops_per_step * args.steps / (tot_ms * 1e-3)

    # ---- the literal configs[1] sweep: 199 successive pairs (latency-bound, reported beside)
    succ_ops = sum(len(succ[ds][0]) for ds in DATASETS) * len(OPS)
    succ_ms = timed(succ, max(args.steps, 20), args.warmup, False)
    succ_val = world * succ_ops * max(args.steps, 20) / (succ_ms * 1e-3)

    # ---- e2e: host roaring_bitmap_t in -> host roaring_bitmap_t out through the C ABI ----------
    e2e = None
    if not args.no_e2e:
        host = {ds: [rb.Bitmap.deserialize(b) for b in blobs[ds]] for ds in DATASETS}
        h2d = d2h = 0
        e2e_chk = [0]

        def e2e_step():
            nonlocal h2d, d2h
            h2d = d2h = 0
            acc = rb.CardinalitySum()
            for ds in DATASETS:
                S = rb.DeviceSet.upload(host[ds]).bind_host()     # H2D inside the timed region; the
                # host inputs stay alive, so pass-through containers are not sent back over PCIe
                h2d += S.payload_bytes
                ia, ib = pairs[ds]
                for op in OPS:
                    r = S.batch(op, S, ia, ib)
                    # queued for the library's background downloader: D2H on its own stream +
                    # host materialisation — every result bitmap is built in the reference layout,
                    # its cardinality read with the host function, then freed (the body of the
                    # reference's benchmark loop, microbenchmarks/bench.cpp:85-96) — while this
                    # thread goes on with the next upload / op
                    r.foreach_async(acc)
                    r.free()
                S.free()
            rb.download_wait()                                    # every result has been consumed
            e2e_chk[0] = acc.value
            d2h = int(rb.api.lib().rb200_last_download_bytes())

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ops_per_step * args.e2e_steps / float(t.item()), "unit": "set-ops/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": args.e2e_steps, "host_threads": host_threads(),
               "checksum_sum_card": e2e_chk[0],
               "api": "rb200_set_upload(host roaring_bitmap_t[]) + rb200_set_bind_host -> per (dataset, op): "
                      "rb200_batch_op + rb200_download_foreach_async -> rb200_download_wait (every result "
                      "materialised as a host roaring_bitmap_t in the reference layout, cardinality read "
                      "on the host, freed; downloads overlap the following uploads and ops)"}

    # ---- e2e, bytes flavour: portable-serialized bitmaps in -> portable-serialized results out
    # (device-side serialization, one D2H per op, no per-container host allocation)
    e2e_ser = None
    if not args.no_e2e:
        ser_bytes = 0

        def e2e_ser_step():
            nonlocal ser_bytes
            ser_bytes = 0
            for ds in DATASETS:
                S = rb.DeviceSet.from_serialized(blobs[ds])        # host parse + H2D
                ia, ib = pairs[ds]
                for op in OPS:
                    r = S.batch(op, S, ia, ib)
                    _buf, _off, _len, release = r.serialize_all(copy=False)
                    ser_bytes += int(rb.api.lib().rb200_last_download_bytes())
                    release()
                    r.free()
                S.free()

        e2e_ser_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_ser_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ser = {"value": world * ops_per_step * args.e2e_steps / float(t.item()), "unit": "set-ops/s",
                   "h2d_bytes_per_step": int(sum(sum(map(len, blobs[ds])) for ds in DATASETS)),
                   "d2h_bytes_per_step": int(ser_bytes),
                   "api": "rb200_set_upload_serialized(portable bytes) -> rb200_batch_op -> "
                          "rb200_set_serialize (portable bytes of every result in pinned host memory)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_compute_items) ------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = stats["algo_bytes"] / (stats["kernel_ms"] * 1e-3) / 1e9 if stats["kernel_ms"] else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("realdata_allpairs", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_compute_items", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": stats["algo_bytes"] / max(stats["launches"], 1),
                "kernel_ms_per_launch": stats["kernel_ms"] / max(stats["launches"], 1),
                "launches_timed": stats["launches"],
                "note": "inputs (0.2-8.7 MB per dataset) are L2-resident within a step by reuse; "
                        "results stream to HBM"}

    cpu = None
    if not args.no_cpu:
        from oracle.refbench import RefBench
        rbn = RefBench()
        T = host_threads()
        samp_ops, samp_t, chk = 0, 0.0, 0
        for ds in DATASETS:
            h = rbn.load(blobs[ds])
            ia, ib = pairs[ds]
            sl = slice(0, len(ia), 4)                      # every 4th pair: ~45k ops, a few seconds
            for op in OPS:
                dt, s = rbn.pairs(h, op, ia[sl], ib[sl], T)
                samp_t += dt
                samp_ops += len(ia[sl])
                chk += s
            rbn.unload(h)
        cpu = {"value": samp_ops / samp_t, "unit": "set-ops/s", "cores": T, "kind": "reference",
               "isa": rbn.isa(),
               "sample": "every 4th of the 19 900 pairs per dataset x and/or/xor x 3 datasets, "
                         "one pass, pairs split over all host threads"}

    line = {
        "metric": METRIC, "value": value, "unit": "set-ops/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16/u64 bitwise",
        "data": "reference realdata fixtures (portable-serialized, run-optimized)",
        "config": {"workload": "realdata_allpairs", "datasets": DATASETS, "ops": OPS,
                   "pairs_per_dataset": int(len(pairs[DATASETS[0]][0])),
                   "set_ops_per_step": ops_per_step, "l2": "flushed between timed steps (256 MB memset)",
                   "batching": "one rb200_batch_op call per (dataset, op); the 9 calls of a step queue back to back, results consumed after"},
        "checksum_sum_card": stats["checksum"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "e2e_serialized": e2e_ser,
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "successive": {"workload": "realdata_successive (configs[1] literal: 199 pairs x 3 ops x 3 datasets)",
                       "value": succ_val, "unit": "set-ops/s", "set_ops_per_step": succ_ops,
                       "ms_per_step": succ_ms / max(args.steps, 20)},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
