#!/usr/bin/env python
"""ncu target: one config-3 density of roaring_bitmap_or_many (200 Zipfian bitmaps x 10^7 values)
or the config-5 shape, a few calls of rb200_or_many after the upload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import croaring_b200 as rb  # noqa: E402
from croaring_b200 import workloads as wl  # noqa: E402

d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 200
rb.init(0)
if d <= 0:      # config 5 shape
    A = wl.zipf_arena(nb, 10 ** 8, None, density_draw=True)
else:
    A = wl.zipf_arena(nb, wl.zipf_universe(10 ** 7, d), 10 ** 7)
S = rb.DeviceSet.from_serialized(A)
for _ in range(reps):
    r = S.or_many()
    print("or_many d=%s: device ms %.3f kernel ms %.3f in GB %.3f" % (
        d, rb.last_device_ms(), rb.api.lib().rb200_last_compute_ms(), S.payload_bytes / 1e9), flush=True)
    r.free()
