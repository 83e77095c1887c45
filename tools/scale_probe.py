#!/usr/bin/env python
"""How k_compute_items scales DOWN: per-(dataset, op) kernel time for every stride-th pair of the
all-pairs batch (stride 1, 2, 4, 8, 16 = what one rank of an N-GPU strong-scaling run holds).
With a -DRB200_PROBE build (RB200_LIB) also the slowest item / warp of each launch in clocks.
Tuning aid; prints one JSON line per (dataset, op)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import croaring_b200 as rb  # noqa: E402

TYPES = {0: "-", 1: "B", 2: "A", 3: "R"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ops", default="and,or")
    ap.add_argument("--datasets", default="census1881,weather_sept_85,wikileaks-noquotes")
    ap.add_argument("--strides", default="1,2,4,8,16")
    a = ap.parse_args()
    rb.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    rb.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    probe = getattr(rb.lib(), "rb200_debug_probe", None) if hasattr(rb.lib(), "rb200_debug_probe") else None
    buf = (ctypes.c_ulonglong * 24)()
    for ds in a.datasets.split(","):
        blobs = rb.load_realdata(ds)
        S = rb.DeviceSet.from_serialized(blobs)
        i, j = np.triu_indices(len(blobs), 1)
        for op in a.ops.split(","):
            row = {"ds": ds, "op": op}
            for st in [int(x) for x in a.strides.split(",")]:
                ia, ib = i[::st].astype(np.uint32), j[::st].astype(np.uint32)
                ks, os_ = [], []
                pr = None
                for rep in range(a.reps + 2):
                    flush.zero_()
                    torch.cuda.synchronize()
                    if probe:
                        probe(buf)
                    r = S.batch(op, S, ia, ib)
                    ms, cms, ab = r.op_stats()
                    if probe:
                        probe(buf)
                        pr = list(buf)
                    r.free()
                    if rep >= 2:
                        ks.append(cms)
                        os_.append(ms)
                cell = {"pairs": len(ia), "kernel_us": round(float(np.median(ks)) * 1e3, 1),
                        "op_us": round(float(np.median(os_)) * 1e3, 1)}
                if pr:
                    tag = pr[0] & 0xFFFFFFFF
                    cell["slowest_item"] = {"clk": pr[0] >> 32, "kind": tag >> 28, "tA": TYPES[(tag >> 26) & 3], "tB": TYPES[(tag >> 24) & 3],
                                            "cA": ((tag >> 12) & 0xFFF) << 5, "cB": (tag & 0xFFF) << 5}
                    cell["slowest_warp"] = {"clk": pr[1] >> 32, "items": pr[1] & 0xFFFFFFFF}
                    cell["fastest_warp"] = {"clk": pr[2] >> 32, "items": pr[2] & 0xFFFFFFFF}
                    cell["items"] = pr[4]
                    cell["avg_item_clk"] = round(pr[3] / max(pr[4], 1))
                    cell["cell_clk_sum"] = pr[5]
                    cell["copy_clk_sum"] = pr[6]
                    cell["warp_clk_sum"] = pr[7]
                    names = ["RUN_ACC", "BR", "AA_ACC", "BA", "BB", "AA", "RUN_IV", "COPY"]
                    cell["classes"] = {names[c]: {"items": pr[16 + c], "avg_clk": round(pr[8 + c] / max(pr[16 + c], 1)),
                                                  "share": round(pr[8 + c] / max(pr[3], 1), 3)} for c in range(8) if pr[16 + c]}
                row[f"s{st}"] = cell
            print(json.dumps(row), flush=True)
        S.free()


if __name__ == "__main__":
    main()
