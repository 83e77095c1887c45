"""ctypes binding of oracle/liboracle.so (the plain-C restatement) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")

OPS = {"and": 0, "or": 1, "xor": 2, "andnot": 3,
       "and_inplace": 4, "or_inplace": 5, "xor_inplace": 6, "andnot_inplace": 7}
MANY = {"or_many": 0, "xor_many": 1}


def build_oracle():
    src = os.path.join(_HERE, "roaring_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self):
        build_oracle()
        L = C.CDLL(ORACLE_SO, mode=os.RTLD_LOCAL)
        self.L = L
        L.oracle_pair_op.restype = C.c_size_t
        L.oracle_pair_op.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                     C.c_char_p, C.c_size_t]
        L.oracle_many_op.restype = C.c_size_t
        L.oracle_many_op.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_char_p),
                                     C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        L.oracle_lazy_fold.restype = C.c_size_t
        L.oracle_lazy_fold.argtypes = [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_char_p),
                                       C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        L.oracle_or_many_heap.restype = C.c_size_t
        L.oracle_or_many_heap.argtypes = [C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                          C.c_char_p, C.c_size_t]
        L.oracle_flip.restype = C.c_size_t
        L.oracle_flip.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t]
        L.oracle_r64_pair_op.restype = C.c_size_t
        L.oracle_r64_pair_op.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                         C.c_char_p, C.c_size_t]
        L.oracle_and_cardinality.restype = C.c_uint64
        L.oracle_and_cardinality.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.oracle_cardinality.restype = C.c_uint64
        L.oracle_cardinality.argtypes = [C.c_char_p, C.c_size_t]

    def op_bytes(self, name: str, a: bytes, b: bytes) -> bytes:
        need = self.L.oracle_pair_op(OPS[name], a, len(a), b, len(b), None, 0)
        if need == C.c_size_t(-1).value:
            raise ValueError("oracle: malformed input")
        buf = C.create_string_buffer(need)
        got = self.L.oracle_pair_op(OPS[name], a, len(a), b, len(b), buf, need)
        assert got == need
        return buf.raw

    def many_bytes(self, name: str, blobs) -> bytes:
        n = len(blobs)
        arr = (C.c_char_p * n)(*blobs)
        lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        need = self.L.oracle_many_op(MANY[name], n, arr, lens, None, 0)
        if need == C.c_size_t(-1).value:
            raise ValueError("oracle: malformed input")
        buf = C.create_string_buffer(need)
        got = self.L.oracle_many_op(MANY[name], n, arr, lens, buf, need)
        assert got == need
        return buf.raw

    def _many(self, fn, head, blobs) -> bytes:
        n = len(blobs)
        arr = (C.c_char_p * n)(*blobs)
        lens = (C.c_size_t * n)(*[len(b) for b in blobs])
        need = fn(*head, n, arr, lens, None, 0)
        if need == C.c_size_t(-1).value:
            raise ValueError("oracle: malformed input")
        buf = C.create_string_buffer(need)
        assert fn(*head, n, arr, lens, buf, need) == need
        return buf.raw

    def lazy_fold_bytes(self, op: str, conv: bool, blobs) -> bytes:
        return self._many(self.L.oracle_lazy_fold, (0 if op == "or" else 1, int(conv)), blobs)

    def or_many_heap_bytes(self, blobs) -> bytes:
        return self._many(self.L.oracle_or_many_heap, (), blobs)

    def r64_op_bytes(self, name: str, a: bytes, b: bytes) -> bytes:
        need = self.L.oracle_r64_pair_op(OPS[name], a, len(a), b, len(b), None, 0)
        if need == C.c_size_t(-1).value:
            raise ValueError("oracle: malformed 64-bit input")
        buf = C.create_string_buffer(need)
        assert self.L.oracle_r64_pair_op(OPS[name], a, len(a), b, len(b), buf, need) == need
        return buf.raw

    def flip_bytes(self, a: bytes, start: int, end: int) -> bytes:
        need = self.L.oracle_flip(a, len(a), start, end, None, 0)
        if need == C.c_size_t(-1).value:
            raise ValueError("oracle: malformed input")
        buf = C.create_string_buffer(need)
        assert self.L.oracle_flip(a, len(a), start, end, buf, need) == need
        return buf.raw

    def and_cardinality(self, a: bytes, b: bytes) -> int:
        return int(self.L.oracle_and_cardinality(a, len(a), b, len(b)))

    def cardinality(self, a: bytes) -> int:
        return int(self.L.oracle_cardinality(a, len(a)))


_o = None


def oracle():
    global _o
    if _o is None:
        _o = Oracle()
    return _o
