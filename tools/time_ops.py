#!/usr/bin/env python
"""Per-(dataset, op) device times of the headline workload for the library RB200_LIB points at
(default: the product build): compute-kernel ms and whole-op ms, median of --reps, L2 flushed.
One JSON line.  Tuning aid: build variants with `python -m croaring_b200.build -DNAME=V --out=...`."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import croaring_b200 as rb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--ops", default="and,or,xor,andnot")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("RB200_LIB", "product")))
    a = ap.parse_args()
    rb.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rb.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"tag": a.tag, "ops": {}}
    tot_k = tot_o = 0.0
    for ds in ["census1881", "weather_sept_85", "wikileaks-noquotes"]:
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        i, j = np.triu_indices(len(blobs), 1)
        ia, ib = i.astype(np.uint32), j.astype(np.uint32)
        for op in a.ops.split(","):
            ks, os_, chk = [], [], 0
            for rep in range(a.reps + 2):
                flush.zero_()
                torch.cuda.synchronize()
                r = S.batch(op, S, ia, ib)
                ms, cms, ab = r.op_stats()
                chk = int(r.cardinalities().sum())
                r.free()
                if rep >= 2:
                    ks.append(cms)
                    os_.append(ms)
            out["ops"][f"{ds}/{op}"] = {"kernel_ms": round(float(np.median(ks)), 4), "op_ms": round(float(np.median(os_)), 4),
                                        "gbs": round(ab / np.median(ks) / 1e6, 1), "sum_card": chk}
            if op in ("and", "or", "xor"):
                tot_k += float(np.median(ks))
                tot_o += float(np.median(os_))
        # the literal 199-successive-pairs call: host enqueue time vs device time of one call
        sa, sb = np.arange(len(blobs) - 1, dtype=np.uint32), np.arange(1, len(blobs), dtype=np.uint32)
        import time
        host_us, dev_ms = [], []
        for rep in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = S.batch("or", S, sa, sb)
            t1 = time.perf_counter()
            ms, cms, ab = r.op_stats()
            r.free()
            if rep >= 5:
                host_us.append((t1 - t0) * 1e6)
                dev_ms.append(ms)
        out["ops"][f"{ds}/successive_or"] = {"host_enqueue_us": round(float(np.median(host_us)), 1),
                                             "device_us": round(float(np.median(dev_ms)) * 1e3, 1)}
        S.free()
    # drop-in single-pair latency (census1881 csv0 AND / OR csv1 through the reference-named symbols)
    import time as _t
    blobs = rb.load_realdata("census1881")
    a, b = rb.Bitmap.deserialize(blobs[0]), rb.Bitmap.deserialize(blobs[1])
    for name, fn in (("and", lambda: a & b), ("or", lambda: a | b)):
        ts = []
        for rep in range(120):
            t0 = _t.perf_counter()
            r = fn()
            ts.append(_t.perf_counter() - t0)
            r.free()
        out["ops"][f"dropin_{name}_us"] = round(float(np.median(ts[20:]) * 1e6), 1)
    out["step_kernel_ms"] = round(tot_k, 4)
    out["step_op_ms"] = round(tot_o, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
