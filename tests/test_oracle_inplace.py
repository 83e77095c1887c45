"""CPU test: the oracle's in-place rules (container_ior full-run conversion, kept full containers)
are pinned against the unmodified reference's roaring_bitmap_*_inplace."""
import numpy as np

from helpers import OPS, synth_blobs


def test_oracle_inplace_vs_reference(O, R):
    blobs = synth_blobs(R, 71, 60, key_space=6, max_keys=7,
                        profiles=["full", "nearfull", "halves", "dense", "bitset", "array", "longruns", "tiny"])
    rng = np.random.default_rng(1)
    for _ in range(500):
        i, j = rng.integers(0, len(blobs), 2)
        for op in OPS:
            assert O.op_bytes(op + "_inplace", blobs[i], blobs[j]) == R.op_inplace_bytes(op, blobs[i], blobs[j]), (op, i, j)
