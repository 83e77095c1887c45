#!/usr/bin/env python
"""Where does the strict e2e step go?  Per (dataset, op): upload / batch / visitor download."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import croaring_b200 as rb
rb.init(0)
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
            t2 = time.perf_counter(); c = r.foreach_sum_cardinality()
            t3 = time.perf_counter(); r.free(); S.free()
            row = (t3 - t0, t1 - t0, t2 - t1, t3 - t2)
            if best is None or row[0] < best[0]:
                best = row
        mb = rb.api.lib().rb200_last_download_bytes() / 1e6
        tot += best[0]
        print(f"{ds:20s} {op:3s}: total {1e3*best[0]:7.2f} ms = upload {1e3*best[1]:.2f} + batch {1e3*best[2]:.2f} + "
              f"foreach {1e3*best[3]:.2f} ({mb:.0f} MB -> {mb/1e3/best[3]:.1f} GB/s)", flush=True)
print(f"sum of best per-op totals: {1e3*tot:.1f} ms per step (threads {os.environ.get('RB200_HOST_THREADS','default 64')})")
