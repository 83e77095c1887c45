"""CPU check (-m "not gpu"): the C examples of INTEGRATION.md compile against include/roaring_b200.h
(names, argument order and types in the documentation follow the header)."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r"""
#include <stdlib.h>
#include <roaring_b200.h>
extern const roaring_bitmap_t **bitmaps;
extern const char *const *bufs, *const *blobs, *const *a64, *const *b64;
extern const size_t *lens, *a64len, *b64len;
extern size_t n, npairs, na, nb;
extern const uint32_t *ia, *ib, *ka, *kb;
extern int rank, nranks, local_gpu;
void consume(roaring_bitmap_t **out, size_t k);
void my_broadcast(void *p, size_t bytes);
"""


def test_integration_examples_compile():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```c\n(.*?)```", text, flags=re.S) if "rb200_" in b and "#ifdef" not in b]
    assert len(blocks) >= 2
    with tempfile.TemporaryDirectory() as d:
        for i, b in enumerate(blocks):
            body = "\n".join(ln for ln in b.splitlines() if not ln.startswith("#include"))
            src = os.path.join(d, f"example{i}.c")
            with open(src, "w") as f:
                f.write(PRELUDE + f"\nvoid example{i}(void) {{\n{body}\n}}\n")
            r = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration",
                                "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
                                "-Wno-unused-variable", "-Wno-unused-but-set-variable",
                                "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
            assert r.returncode == 0, f"INTEGRATION.md example {i} does not compile:\n{r.stderr[-3000:]}"


def test_header_coexists_with_the_reference_headers():
    """include/roaring_b200.h next to the reference's own headers, in either include order, as C and as
    C++: the drop-in prototypes must be compatible declarations of the reference's (a conflicting
    signature is a compile error) and the layout types must not be defined twice."""
    ref_inc = "/root/reference/include"
    if not os.path.isdir(ref_inc):
        import pytest
        pytest.skip("needs the reference headers (build container)")
    orders = [("#include <roaring/roaring.h>\n#include <roaring/roaring64.h>\n#include <roaring_b200.h>\n"),
              ("#include <roaring_b200.h>\n#include <roaring/roaring.h>\n#include <roaring/roaring64.h>\n")]
    with tempfile.TemporaryDirectory() as d:
        for i, inc in enumerate(orders):
            for cc, std, ext in (("gcc", "-std=c11", "c"), ("g++", "-std=c++17", "cpp")):
                src = os.path.join(d, f"order{i}.{ext}")
                with open(src, "w") as f:
                    f.write(inc + "int main(void) { roaring_bitmap_t *r = 0; (void)r; return 0; }\n")
                r = subprocess.run([cc, std, "-fsyntax-only", "-Wall", "-I", ref_inc, "-I", os.path.join(ROOT, "include"), src],
                                   capture_output=True, text=True)
                assert r.returncode == 0, f"{cc} include order {i}:\n{r.stderr[-3000:]}"
        # and alone, without the reference anywhere on the include path
        src = os.path.join(d, "alone.c")
        with open(src, "w") as f:
            f.write("#include <roaring_b200.h>\nint main(void) { roaring_bitmap_t r; (void)r; return 0; }\n")
        r = subprocess.run(["gcc", "-std=c11", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), src],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
