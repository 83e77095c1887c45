#!/bin/bash
# Round-2 GPU call W (1 minute left): multi-chunk window index parity only (no pytest / torch import).
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
RB200_OR_MANY_INDEX=window timeout 50 python - > gpurun_out/w.log 2>&1 <<'PY'
import sys, os, time
t0 = time.time()
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.getcwd() + "/tests")
import numpy as np
import croaring_b200 as rb
from helpers import synth_blobs
from oracle.refbind import ref
R = ref()
rb.init(0)
blobs = synth_blobs(R, 123, 700, key_space=300, max_keys=80, profiles=["array", "tiny", "tiny", "tiny", "longruns", "bitset", "full", "array"])
S = rb.DeviceSet.from_serialized(blobs)
out = S.or_many().download(0)
print("700-way", out.serialize() == R.many_bytes("or_many", blobs), time.time() - t0, flush=True)
idx = np.arange(0, 700, 3, dtype=np.uint32)
out = S.or_many(idx).download(0)
print("subset", out.serialize() == R.many_bytes("or_many", [blobs[i] for i in idx]), time.time() - t0, flush=True)
blobs = synth_blobs(R, 7, 1500, key_space=2000, max_keys=200, profiles=["array", "tiny", "bitset", "longruns"])
S = rb.DeviceSet.from_serialized(blobs)
out = S.or_many().download(0)
print("1500-way", out.serialize() == R.many_bytes("or_many", blobs), time.time() - t0, flush=True)
PY
cat gpurun_out/w.log
