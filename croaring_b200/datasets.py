"""Fixture / workload I/O shared by tests and bench (pure Python, no reference, no oracle).

.rbset file: b"RBSET1\\0\\0" | u32 n | n x u32 byte-length | n portable-serialized bitmaps,
xz-compressed.  The committed real-data sets under tests/golden/realdata/ were written by
tests/golden/make_golden.py with the reference's own roaring_bitmap_portable_serialize.
"""
import lzma
import os
import struct

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REALDATA_DIR = os.path.join(_ROOT, "tests", "golden", "realdata")
MAGIC = b"RBSET1\0\0"


def write_rbset(path, blobs, preset=9 | lzma.PRESET_EXTREME):
    raw = MAGIC + struct.pack("<I", len(blobs))
    raw += b"".join(struct.pack("<I", len(b)) for b in blobs) + b"".join(blobs)
    with open(path, "wb") as f:
        f.write(lzma.compress(raw, preset=preset))


def read_rbset(path):
    with open(path, "rb") as f:
        raw = lzma.decompress(f.read())
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not an .rbset file")
    n, = struct.unpack_from("<I", raw, 8)
    lens = struct.unpack_from(f"<{n}I", raw, 12)
    off = 12 + 4 * n
    out = []
    for ln in lens:
        out.append(raw[off:off + ln])
        off += ln
    return out


def load_realdata(name):
    """The 200 portable-serialized bitmaps of a reference real-data set (run-optimized)."""
    return read_rbset(os.path.join(REALDATA_DIR, name + ".rbset.xz"))
