#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Runs only in the build container (needs /root/reference for the raw datasets and
oracle/_ref/libroaring_ref.so built by `make -C oracle`).  Outputs (committed):

  tests/golden/realdata/<dataset>.rbset.xz
      the 200 bitmaps of a reference real-data set, built exactly like the reference's
      microbenchmark loader (/root/reference/microbenchmarks/bench.h:286-313 ->
      roaring_bitmap_of_ptr + run_optimize + shrink_to_fit; files in scandir/alphasort
      order, /root/reference/benchmarks/numbersfromtextfiles.h:105-140) and written with the
      reference's own roaring_bitmap_portable_serialize ("identical serialized inputs").
  tests/golden/realdata_golden.json
      golden values computed BY THE REFERENCE on those inputs: sum of result cardinalities
      of the 199 successive pairs for and/or/xor/andnot (/root/reference/microbenchmarks/
      bench.cpp:85-96,196-207), sha256 over the concatenated portable serialisations of the
      199 results (pins result container TYPES too), or_many / xor_many cardinality + sha256.
      Same again for the twins without run containers (remove_run_compression), as
      /root/reference/tests/realdata_unit.c:796-806 exercises.

File format .rbset:  b"RBSET1\\0\\0" | u32 n | n x u32 byte-length | blobs.
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
from croaring_b200.datasets import read_rbset, write_rbset  # noqa: E402

REALDATA = "/root/reference/benchmarks/realdata"
DATASETS = ["census1881", "weather_sept_85", "wikileaks-noquotes",
            "census1881_srt", "wikileaks-noquotes_srt", "uscensus2000"]


def read_ints(path):
    with open(path, "rb") as f:
        txt = f.read()
    return np.array([int(t) for t in txt.replace(b"\n", b",").split(b",") if t.strip()],
                    dtype=np.uint32)


def golden_for(R, bms):
    g = {}
    n = len(bms)
    for op in ("and", "or", "xor", "andnot"):
        h = hashlib.sha256()
        tot = 0
        for i in range(n - 1):
            r = R.op(op, bms[i], bms[i + 1])
            ok, why = R.validate(r)
            assert ok, why
            tot += R.card(r)
            h.update(R.serialize(r))
            R.free(r)
        g[op] = {"sum_card": tot, "sha256": h.hexdigest()}
    g["and_cardinality"] = sum(
        int(R.L.roaring_bitmap_and_cardinality(bms[i], bms[i + 1])) for i in range(n - 1))
    for name in ("or_many", "xor_many"):
        r = R.many(name, bms)
        g[name] = {"card": R.card(r), "sha256": hashlib.sha256(R.serialize(r)).hexdigest()}
        R.free(r)
    return g


def main():
    R = ref()
    os.makedirs(os.path.join(HERE, "realdata"), exist_ok=True)
    golden = {}
    for ds in DATASETS:
        d = os.path.join(REALDATA, ds)
        files = sorted(f for f in os.listdir(d) if f.endswith(".txt"))  # alphasort == strcmp
        bms = [R.from_values(read_ints(os.path.join(d, f)), run_optimize=True) for f in files]
        blobs = [R.serialize(b) for b in bms]
        write_rbset(os.path.join(HERE, "realdata", ds + ".rbset.xz"), blobs)
        assert read_rbset(os.path.join(HERE, "realdata", ds + ".rbset.xz")) == blobs
        g = {"n": len(bms), "portable_bytes": sum(map(len, blobs)),
             "first_files": files[:2], "run_optimized": golden_for(R, bms)}
        for b in bms:
            R.L.roaring_bitmap_remove_run_compression(b)
        g["no_runs"] = golden_for(R, bms)
        golden[ds] = g
        for b in bms:
            R.free(b)
        print(ds, g["portable_bytes"], g["run_optimized"]["and"]["sum_card"],
              g["run_optimized"]["or_many"]["card"], flush=True)
    with open(os.path.join(HERE, "realdata_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
