"""ctypes binding of oracle/_ref/libroaring_ref.so — TEST INFRASTRUCTURE ONLY.

`libroaring_ref.so` is the UNMODIFIED reference (CRoaring 5.1.0) compiled from the sources
under /root/reference by oracle/Makefile.  It is the ground truth that (a) pins the plain-C
restatement in oracle/roaring_oracle.c and (b) is the differential checker for the CUDA path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (croaring_b200/) never does.

Signatures follow /root/reference/include/roaring/roaring.h (line numbers in comments).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libroaring_ref.so")

c_bitmap_p = C.c_void_p


class RefLib:
    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(path, mode=os.RTLD_LOCAL)
        self.L = L

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)
            return f

        P = c_bitmap_p
        sig("roaring_bitmap_create_with_capacity", P, C.c_uint32)           # roaring.h:54
        sig("roaring_bitmap_of_ptr", P, C.c_size_t, C.c_void_p)             # roaring.h:97
        sig("roaring_bitmap_add_many", None, P, C.c_size_t, C.c_void_p)
        sig("roaring_bitmap_add_range_closed", None, P, C.c_uint32, C.c_uint32)
        sig("roaring_bitmap_copy", P, P)
        sig("roaring_bitmap_free", None, P)                                 # roaring.h:365
        sig("roaring_bitmap_run_optimize", C.c_bool, P)
        sig("roaring_bitmap_remove_run_compression", C.c_bool, P)
        sig("roaring_bitmap_shrink_to_fit", C.c_size_t, P)
        sig("roaring_bitmap_get_cardinality", C.c_uint64, P)
        sig("roaring_bitmap_portable_size_in_bytes", C.c_size_t, P)
        sig("roaring_bitmap_portable_serialize", C.c_size_t, P, C.c_char_p)
        sig("roaring_bitmap_portable_deserialize_safe", P, C.c_char_p, C.c_size_t)
        sig("roaring_bitmap_internal_validate", C.c_bool, P, C.POINTER(C.c_char_p))
        sig("roaring_bitmap_equals", C.c_bool, P, P)
        sig("roaring_bitmap_is_subset", C.c_bool, P, P)                     # roaring.h:905
        sig("roaring_bitmap_is_strict_subset", C.c_bool, P, P)              # roaring.h:912
        sig("roaring_bitmap_to_uint32_array", None, P, C.c_void_p)
        sig("roaring_bitmap_set_copy_on_write", None, P, C.c_bool)
        for op in ("and", "or", "xor", "andnot"):                          # roaring.h:225,288,320,342
            sig(f"roaring_bitmap_{op}", P, P, P)
            sig(f"roaring_bitmap_{op}_inplace", None, P, P)
        sig("roaring_bitmap_or_many", P, C.c_size_t, C.POINTER(P))          # roaring.h:304
        sig("roaring_bitmap_or_many_heap", P, C.c_uint32, C.POINTER(P))     # roaring.h:312
        sig("roaring_bitmap_xor_many", P, C.c_size_t, C.POINTER(P))         # roaring.h:334
        for op in ("and", "or", "xor", "andnot"):                          # roaring.h:231,258-271
            sig(f"roaring_bitmap_{op}_cardinality", C.c_uint64, P, P)
        sig("roaring_bitmap_jaccard_index", C.c_double, P, P)               # roaring.h:252
        sig("roaring_bitmap_intersect", C.c_bool, P, P)                     # roaring.h:237
        sig("roaring_bitmap_statistics", None, P, C.c_void_p)
        sig("roaring_bitmap_frozen_size_in_bytes", C.c_size_t, P)           # roaring.h:831
        sig("roaring_bitmap_frozen_serialize", None, P, C.c_void_p)         # roaring.h:846
        sig("roaring_bitmap_frozen_view", P, C.c_void_p, C.c_size_t)        # roaring.h:864
        V = C.c_void_p                                                      # roaring64_bitmap_t*
        sig("roaring64_bitmap_of_ptr", V, C.c_size_t, C.c_void_p)           # roaring64.h:92
        sig("roaring64_bitmap_free", None, V)                               # roaring64.h:69
        sig("roaring64_bitmap_run_optimize", C.c_bool, V)                   # roaring64.h:364
        sig("roaring64_bitmap_get_cardinality", C.c_uint64, V)              # roaring64.h:326
        sig("roaring64_bitmap_internal_validate", C.c_bool, V, C.POINTER(C.c_char_p))
        sig("roaring64_bitmap_portable_size_in_bytes", C.c_size_t, V)
        sig("roaring64_bitmap_portable_serialize", C.c_size_t, V, C.c_char_p)
        sig("roaring64_bitmap_portable_deserialize_safe", V, C.c_char_p, C.c_size_t)
        for op in ("and", "or", "xor", "andnot"):                          # roaring64.h:423-530
            sig(f"roaring64_bitmap_{op}", V, V, V)
        sig("roaring64_bitmap_and_cardinality", C.c_uint64, V, V)           # roaring64.h:429
        sig("roaring_bitmap_flip", P, P, C.c_uint64, C.c_uint64)            # roaring.h:986
        sig("roaring_bitmap_flip_inplace", None, P, C.c_uint64, C.c_uint64) # roaring.h:1004
        sig("roaring_bitmap_lazy_or", P, P, P, C.c_bool)                    # roaring.h:932
        sig("roaring_bitmap_lazy_or_inplace", None, P, P, C.c_bool)         # roaring.h:943
        sig("roaring_bitmap_lazy_xor", P, P, P)                             # roaring.h:963
        sig("roaring_bitmap_lazy_xor_inplace", None, P, P)                  # roaring.h:970
        sig("roaring_bitmap_repair_after_lazy", None, P)                    # roaring.h:952

    # ---- helpers -------------------------------------------------------------
    def from_values(self, vals, run_optimize=True):
        a = np.ascontiguousarray(vals, dtype=np.uint32)
        r = self.L.roaring_bitmap_of_ptr(a.size, a.ctypes.data)
        if run_optimize:
            self.L.roaring_bitmap_run_optimize(r)
        self.L.roaring_bitmap_shrink_to_fit(r)
        return r

    def serialize(self, r) -> bytes:
        n = self.L.roaring_bitmap_portable_size_in_bytes(r)
        buf = C.create_string_buffer(n)
        m = self.L.roaring_bitmap_portable_serialize(r, buf)
        assert m == n
        return buf.raw

    def deserialize(self, b: bytes):
        r = self.L.roaring_bitmap_portable_deserialize_safe(b, len(b))
        if not r:
            raise ValueError("reference refused to deserialize")
        return r

    def free(self, r):
        self.L.roaring_bitmap_free(r)

    def validate(self, r):
        reason = C.c_char_p()
        ok = self.L.roaring_bitmap_internal_validate(r, C.byref(reason))
        return bool(ok), (reason.value.decode() if reason.value else "")

    def card(self, r):
        return int(self.L.roaring_bitmap_get_cardinality(r))

    def to_array(self, r):
        n = self.card(r)
        out = np.empty(n, dtype=np.uint32)
        if n:
            self.L.roaring_bitmap_to_uint32_array(r, out.ctypes.data)
        return out

    def op(self, name, a, b):
        return getattr(self.L, f"roaring_bitmap_{name}")(a, b)

    def many(self, name, rs):
        arr = (c_bitmap_p * len(rs))(*rs)
        if name == "or_many_heap":
            return self.L.roaring_bitmap_or_many_heap(len(rs), arr)
        return getattr(self.L, f"roaring_bitmap_{name}")(len(rs), arr)

    # bytes -> bytes convenience used by the oracle-pinning tests
    def op_bytes(self, name, a: bytes, b: bytes) -> bytes:
        ra, rb = self.deserialize(a), self.deserialize(b)
        r = self.op(name, ra, rb)
        out = self.serialize(r)
        for x in (ra, rb, r):
            self.free(x)
        return out

    def op_inplace_bytes(self, name, a: bytes, b: bytes) -> bytes:
        """roaring_bitmap_<name>_inplace(x1, x2) on fresh deserializations; returns x1's bytes."""
        ra, rb = self.deserialize(a), self.deserialize(b)
        getattr(self.L, f"roaring_bitmap_{name}_inplace")(ra, rb)
        out = self.serialize(ra)
        ok, why = self.validate(ra)
        assert ok, why
        self.free(ra)
        self.free(rb)
        return out

    # ---- 64-bit bitmaps (roaring64.h), through their portable format
    def r64_from_values(self, vals, run_optimize=True) -> bytes:
        v = np.ascontiguousarray(vals, dtype=np.uint64)
        r = self.L.roaring64_bitmap_of_ptr(v.size, v.ctypes.data)
        if run_optimize:
            self.L.roaring64_bitmap_run_optimize(r)
        b = self.r64_serialize(r)
        self.L.roaring64_bitmap_free(r)
        return b

    def r64_serialize(self, r) -> bytes:
        n = self.L.roaring64_bitmap_portable_size_in_bytes(r)
        buf = C.create_string_buffer(n)
        assert self.L.roaring64_bitmap_portable_serialize(r, buf) == n
        return buf.raw[:n]

    def r64_deserialize(self, b: bytes):
        r = self.L.roaring64_bitmap_portable_deserialize_safe(b, len(b))
        assert r, "reference refused a 64-bit blob"
        return r

    def r64_op_bytes(self, name: str, a: bytes, b: bytes) -> bytes:
        ra, rb = self.r64_deserialize(a), self.r64_deserialize(b)
        r = getattr(self.L, f"roaring64_bitmap_{name}")(ra, rb)
        why = C.c_char_p()
        assert self.L.roaring64_bitmap_internal_validate(r, C.byref(why)), why.value
        out = self.r64_serialize(r)
        for x in (ra, rb, r):
            self.L.roaring64_bitmap_free(x)
        return out

    def r64_and_cardinality(self, a: bytes, b: bytes) -> int:
        ra, rb = self.r64_deserialize(a), self.r64_deserialize(b)
        v = int(self.L.roaring64_bitmap_and_cardinality(ra, rb))
        self.L.roaring64_bitmap_free(ra)
        self.L.roaring64_bitmap_free(rb)
        return v

    def flip_bytes(self, blob: bytes, start: int, end: int, inplace=False) -> bytes:
        r = self.deserialize(blob)
        if inplace:
            self.L.roaring_bitmap_flip_inplace(r, start, end)
            out = r
        else:
            out = self.L.roaring_bitmap_flip(r, start, end)
        ok, why = self.validate(out)
        assert ok, why
        b = self.serialize(out)
        if out != r:
            self.free(out)
        self.free(r)
        return b

    def frozen_bytes(self, blob: bytes) -> bytes:
        """roaring_bitmap_frozen_serialize of the bitmap held in a portable blob."""
        r = self.deserialize(blob)
        n = self.L.roaring_bitmap_frozen_size_in_bytes(r)
        buf = C.create_string_buffer(n)
        self.L.roaring_bitmap_frozen_serialize(r, buf)
        self.free(r)
        return buf.raw[:n]

    def lazy_fold_bytes(self, op: str, conv: bool, blobs) -> bytes:
        """repair_after_lazy(lazy_<op>(x0, x1) then lazy_<op>_inplace(acc, xi) for i >= 2)."""
        rs = [self.deserialize(b) for b in blobs]
        L = self.L
        if len(rs) == 1:
            acc = L.roaring_bitmap_copy(rs[0])
        elif op == "or":
            acc = L.roaring_bitmap_lazy_or(rs[0], rs[1], conv)
            for x in rs[2:]:
                L.roaring_bitmap_lazy_or_inplace(acc, x, conv)
        else:
            acc = L.roaring_bitmap_lazy_xor(rs[0], rs[1])
            for x in rs[2:]:
                L.roaring_bitmap_lazy_xor_inplace(acc, x)
        L.roaring_bitmap_repair_after_lazy(acc)
        ok, why = self.validate(acc)
        assert ok, why
        out = self.serialize(acc)
        for x in rs + [acc]:
            self.free(x)
        return out

    def many_bytes(self, name, blobs) -> bytes:
        rs = [self.deserialize(b) for b in blobs]
        r = self.many(name, rs)
        out = self.serialize(r)
        for x in rs + [r]:
            self.free(x)
        return out


_lib = None


def ref():
    global _lib
    if _lib is None:
        _lib = RefLib()
    return _lib
