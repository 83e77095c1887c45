#!/bin/bash
# Round-2 GPU call K: why k_compute_items does not scale down (strong scaling floor).
mkdir -p gpurun_out
timeout 200 python tools/scale_probe.py > gpurun_out/scale_product.jsonl 2> gpurun_out/scale_product.err
RB200_LIB=$PWD/croaring_b200/_probe.so timeout 200 python tools/scale_probe.py --strides 1,8 > gpurun_out/scale_probe.jsonl 2> gpurun_out/scale_probe.err
RB200_LIB=$PWD/croaring_b200/_t1.so timeout 200 python tools/scale_probe.py --strides 1,4,8,16 > gpurun_out/scale_t1.jsonl 2> gpurun_out/scale_t1.err
RB200_ORDER_MIN=4000000000 timeout 200 python tools/scale_probe.py --strides 4,8,16 > gpurun_out/scale_noorder.jsonl 2> gpurun_out/scale_noorder.err
tail -3 gpurun_out/scale_*.err
cat gpurun_out/scale_product.jsonl | cut -c1-600
