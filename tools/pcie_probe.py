#!/usr/bin/env python
"""PCIe / NUMA probe: pinned D2H and H2D bandwidth with the staging buffer placed on the GPU's
NUMA node vs the other socket (placement = where the allocating thread runs)."""
import glob, os, sys
import torch

def cpulist(s):
    out = []
    for tok in s.strip().split(","):
        if "-" in tok:
            a, b = tok.split("-"); out += list(range(int(a), int(b) + 1))
        elif tok:
            out.append(int(tok))
    return out

torch.cuda.init()
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
import ctypes
rt = ctypes.CDLL("libcudart.so.12") if False else None
busid = None
for d in glob.glob("/sys/bus/pci/devices/*"):
    try:
        if open(d + "/vendor").read().strip() == "0x10de" and open(d + "/class").read().startswith("0x0302"):
            busid = d; break
    except OSError:
        pass
print("gpu pci dir:", busid)
local = []
if busid:
    print("numa_node:", open(busid + "/numa_node").read().strip(), "local_cpulist:", open(busid + "/local_cpulist").read().strip())
    local = cpulist(open(busid + "/local_cpulist").read())
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(os.path.basename(n), open(n + "/cpulist").read().strip())
allc = sorted(os.sched_getaffinity(0))
remote = [c for c in allc if c not in local]
print("affinity:", len(allc), "cpus; local", len(local), "remote", len(remote))

def bw(buf, dev, direction):
    CH = 64 << 20
    n = buf.numel() // CH
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0
    for rep in range(3):
        torch.cuda.synchronize(); s.record()
        for k in range(n):
            if direction == "d2h":
                buf[k * CH:(k + 1) * CH].copy_(dev[k * CH:(k + 1) * CH], non_blocking=True)
            else:
                dev[k * CH:(k + 1) * CH].copy_(buf[k * CH:(k + 1) * CH], non_blocking=True)
        e.record(); torch.cuda.synchronize()
        best = max(best, n * CH / (s.elapsed_time(e) * 1e-3) / 1e9)
    return best

dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, cpus in (("unbound", allc), ("local", local or allc), ("remote", remote or allc)):
    os.sched_setaffinity(0, cpus)
    buf = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); buf.zero_()
    os.sched_setaffinity(0, allc)
    print(f"{name:8s}: D2H {bw(buf, dev, 'd2h'):.1f} GB/s   H2D {bw(buf, dev, 'h2d'):.1f} GB/s", flush=True)
    del buf
