#!/usr/bin/env python
"""Where does the e2e step go?  Per (dataset, op): upload / batch / streaming download+free."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import croaring_b200 as rb
rb.init(0)
# raw pinned D2H bandwidth for reference
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(x); torch.cuda.synchronize()
    print(f"raw D2H 1 GiB pinned: {1.0737/(time.perf_counter()-t0):.1f} GB/s")
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tot = 0.0
for ds in ["census1881", "weather_sept_85", "wikileaks-noquotes"]:
    blobs = rb.load_realdata(ds)
    host = [rb.Bitmap.deserialize(b) for b in blobs]
    i, j = np.triu_indices(len(blobs), 1)
    ia, ib = i.astype(np.uint32), j.astype(np.uint32)
    for op in ["and", "or", "xor"]:
        best = None
        for rep in range(3):
            t0 = time.perf_counter(); S = rb.DeviceSet.upload(host)
            t1 = time.perf_counter(); r = S.batch(op, S, ia, ib)
            t2 = time.perf_counter(); tfree = 0.0; nch = 0
            for arr, n in r.download_stream(chunk):
                tf = time.perf_counter(); rb.DeviceSet.free_raw(arr, n); tfree += time.perf_counter() - tf; nch += 1
            t3 = time.perf_counter(); r.free(); S.free()
            t4 = time.perf_counter()
            row = (t4 - t0, t1 - t0, t2 - t1, t3 - t2, tfree, nch)
            if best is None or row[0] < best[0]:
                best = row
        mb = rb.api.lib().rb200_last_download_bytes() / 1e6
        tot += best[0]
        print(f"{ds:20s} {op:3s}: total {1e3*best[0]:7.2f} ms = upload {1e3*best[1]:.2f} + batch {1e3*best[2]:.2f} + "
              f"stream {1e3*best[3]:.2f} (free {1e3*best[4]:.2f}, {best[5]} chunks, {mb:.0f} MB -> {mb/1e3/best[3]:.1f} GB/s)", flush=True)
print(f"sum of best per-op totals: {1e3*tot:.1f} ms per step")
