"""C-level proof that the library "drops into callers unchanged" (VERDICT r1 item 5): a plain C
program written against the REFERENCE headers (tests/c/dropin_caller.c, the loops of
microbenchmarks/bench.cpp:85-96 and tests/realdata_unit.c:323-446) is linked
  (a) against the reference alone,
  (b) with -lroaring_b200 ahead of -lroaring_ref on the link line,
and (a) is also run under LD_PRELOAD=libroaring_b200.so.  All three must print identical
checksums (cardinalities + hashes of the reference's own serialisation of every result), every
result passes roaring_bitmap_internal_validate and is released by the reference's
roaring_bitmap_free, also with every allocation routed through roaring_init_memory_hook."""
import os
import struct
import subprocess

import pytest

import croaring_b200 as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CALLER_REF = os.path.join(REFDIR, "dropin_caller_ref")
CALLER_B200 = os.path.join(REFDIR, "dropin_caller_b200")


def _write_inputs(path):
    blobs = rb.load_realdata("census1881")[:14] + rb.load_realdata("weather_sept_85")[:10] \
        + rb.load_realdata("wikileaks-noquotes")[:12]
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(blobs)))
        f.write(struct.pack(f"<{len(blobs)}I", *map(len, blobs)))
        for b in blobs:
            f.write(b)


def _run(exe, inp, *args, preload=None):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    p = subprocess.run([exe, inp, *args], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, (exe, p.returncode, p.stdout[-400:], p.stderr[-400:])
    launches = int(p.stderr.strip().split("kernel_launches=")[-1].split()[0])
    return p.stdout, launches


def test_callers_built_and_reference_alone_runs(tmp_path):
    """CPU part: the binaries exist, resolve their libraries, and the reference-only build runs."""
    assert os.path.exists(CALLER_REF) and os.path.exists(CALLER_B200), "run `make -C oracle` where /root/reference exists"
    ldd = subprocess.check_output(["ldd", CALLER_B200], text=True)
    assert "libroaring_b200.so" in ldd and "libroaring_ref.so" in ldd and "not found" not in ldd
    inp = str(tmp_path / "in.bin")
    _write_inputs(inp)
    out, launches = _run(CALLER_REF, inp, "hook")
    assert launches == -1 and "hook outstanding=0" in out


@pytest.mark.gpu
def test_c_caller_link_order_and_preload(tmp_path):
    inp = str(tmp_path / "in.bin")
    _write_inputs(inp)
    ref_out, ref_l = _run(CALLER_REF, inp)
    assert ref_l == -1                                    # the reference alone never touches the GPU
    b200_out, b200_l = _run(CALLER_B200, inp)
    assert b200_l > 100, "link order did not route the hot path to libroaring_b200.so"
    assert b200_out == ref_out
    pre_out, pre_l = _run(CALLER_REF, inp, preload=rb.api.LIB_PATH)
    assert pre_l > 100 and pre_out == ref_out
    # allocator contract: everything we hand out comes from (and returns to) the user's hooks
    hook_ref, _ = _run(CALLER_REF, inp, "hook")
    hook_b200, _ = _run(CALLER_B200, inp, "hook")
    assert hook_b200 == hook_ref and "hook outstanding=0" in hook_b200
