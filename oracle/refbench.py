"""ctypes binding of oracle/_ref/libref_bench.so — times the UNMODIFIED reference on host cores.
TEST/BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline leg and --impl reference arm)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_bench.so")
OPS = {"and": 0, "or": 1, "xor": 2, "andnot": 3, "and_cardinality": 4}


class RefBench:
    def __init__(self):
        if not os.path.exists(SO):
            raise FileNotFoundError(f"{SO} missing: run `make -C oracle` where /root/reference exists")
        # the driver links libroaring_ref.so through rpath=$ORIGIN
        L = C.CDLL(SO, mode=os.RTLD_LOCAL)
        L.refbench_load.restype = C.c_void_p
        L.refbench_load.argtypes = [C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
        L.refbench_unload.restype = None
        L.refbench_unload.argtypes = [C.c_void_p, C.c_size_t]
        L.refbench_pairs.restype = C.c_double
        L.refbench_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.c_int, C.POINTER(C.c_uint64)]
        L.refbench_or_many.restype = C.c_double
        L.refbench_or_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                       C.POINTER(C.c_uint64)]
        L.refbench_hardware_support.restype = C.c_int
        L.refbench_load_mt.restype = C.c_void_p
        L.refbench_load_mt.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.refbench_warm_pool.restype = None
        L.refbench_warm_pool.argtypes = [C.c_int]
        L.refbench_or_many_bytes.restype = C.c_double
        L.refbench_or_many_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.refbench_free.restype = None
        L.refbench_free.argtypes = [C.c_void_p]
        self.L = L

    def load(self, blobs, nthreads=1):
        """blobs: list of bytes or a croaring_b200.workloads.BlobArena (pointers used in place)."""
        n = len(blobs)
        if hasattr(blobs, "ptrs"):
            arr, lens = blobs.ptrs, blobs.lens
        else:
            arr = (C.c_char_p * n)(*blobs)
            lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        h = self.L.refbench_load_mt(n, arr, lens, int(nthreads))
        if not h:
            raise ValueError("reference refused an input")
        return h, n

    def warm_pool(self, nthreads):
        self.L.refbench_warm_pool(int(nthreads))

    def or_many_bytes(self, handle, idx=None):
        """(seconds of the or_many call, portable bytes of its result, cardinality)."""
        if idx is None:
            ip, n = None, handle[1]
        else:
            idx = np.ascontiguousarray(idx, dtype=np.uint32)
            ip, n = idx.ctypes.data, idx.size
        out, ln, c = C.c_void_p(), C.c_size_t(), C.c_uint64()
        dt = self.L.refbench_or_many_bytes(handle[0], ip, n, C.byref(out), C.byref(ln), C.byref(c))
        b = C.string_at(out.value, ln.value)
        self.L.refbench_free(out)
        return float(dt), b, int(c.value)

    def unload(self, handle):
        self.L.refbench_unload(handle[0], handle[1])

    def pairs(self, handle, op, ia, ib, nthreads=1):
        ia = np.ascontiguousarray(ia, dtype=np.uint32)
        ib = np.ascontiguousarray(ib, dtype=np.uint32)
        s = C.c_uint64()
        dt = self.L.refbench_pairs(handle[0], OPS[op], ia.ctypes.data, ib.ctypes.data, ia.size,
                                   int(nthreads), C.byref(s))
        return float(dt), int(s.value)

    def or_many(self, handle, idx, reps=1):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        c = C.c_uint64()
        dt = self.L.refbench_or_many(handle[0], idx.ctypes.data, idx.size, int(reps), C.byref(c))
        return float(dt), int(c.value)

    def isa(self):
        v = self.L.refbench_hardware_support()
        return "avx512" if v & 2 else ("avx2" if v & 1 else "scalar")
