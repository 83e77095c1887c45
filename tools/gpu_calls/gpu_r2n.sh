#!/bin/bash
# Round-2 GPU call N: where k_compute_items' clocks go, class by class (probe build).
mkdir -p gpurun_out
RB200_LIB=$PWD/croaring_b200/_probe.so timeout 300 python tools/scale_probe.py --strides 1 --ops and,or,xor --reps 3 > gpurun_out/scale_classes.jsonl 2> gpurun_out/scale_classes.err
tail -n 3 gpurun_out/scale_classes.err
