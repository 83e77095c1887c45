#!/usr/bin/env python
"""Golden values of the TIMED bench workload (bench.py `realdata_allpairs`, SURVEY.md §8(d)
config 2b), computed BY THE UNMODIFIED REFERENCE (oracle/_ref) on the committed fixtures:

  tests/golden/allpairs_golden.json
      per dataset (census1881, weather_sept_85, wikileaks-noquotes) and op (and / or / xor /
      andnot), over ALL 19 900 unordered pairs (i < j, row-major = numpy.triu_indices order):
      sum of result cardinalities, sum of roaring_bitmap_portable_size_in_bytes, and sha256 over
      the concatenated portable serialisations of the 19 900 results (pins container TYPES);
      plus sum of roaring_bitmap_and_cardinality.

bench.py compares its device checksum with `sum_card` (and with the reference arm run beside it);
tests/test_gpu_allpairs.py compares the sha256 of the device-serialized results.
Needs oracle/_ref/libroaring_ref.so (`make -C oracle`) — not /root/reference.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refbind import ref  # noqa: E402
from croaring_b200.datasets import load_realdata  # noqa: E402

DATASETS = ["census1881", "weather_sept_85", "wikileaks-noquotes"]


def main():
    R = ref()
    out = {}
    for ds in DATASETS:
        bms = [R.deserialize(b) for b in load_realdata(ds)]
        n = len(bms)
        ia, ib = np.triu_indices(n, 1)
        g = {"n": n, "pairs": int(len(ia))}
        for op in ("and", "or", "xor", "andnot"):
            h = hashlib.sha256()
            tot = size = 0
            for i, j in zip(ia.tolist(), ib.tolist()):
                r = R.op(op, bms[i], bms[j])
                tot += R.card(r)
                b = R.serialize(r)
                size += len(b)
                h.update(b)
                R.free(r)
            g[op] = {"sum_card": tot, "sum_portable_bytes": size, "sha256": h.hexdigest()}
        g["and_cardinality"] = sum(int(R.L.roaring_bitmap_and_cardinality(bms[i], bms[j]))
                                   for i, j in zip(ia.tolist(), ib.tolist()))
        assert g["and_cardinality"] == g["and"]["sum_card"]
        out[ds] = g
        for b in bms:
            R.free(b)
        print(ds, {op: g[op]["sum_card"] for op in ("and", "or", "xor", "andnot")}, flush=True)
    out["bench_checksum_and_or_xor"] = sum(out[ds][op]["sum_card"] for ds in DATASETS for op in ("and", "or", "xor"))
    with open(os.path.join(HERE, "allpairs_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("bench checksum", out["bench_checksum_and_or_xor"])


if __name__ == "__main__":
    main()
