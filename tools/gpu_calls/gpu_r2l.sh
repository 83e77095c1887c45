#!/bin/bash
# Round-2 GPU call L: (1) why k_compute_items does not scale down; (2) or_many window index + flat arrays.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_many_index.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_properties.py tests/test_gpu_xor_many.py -m gpu -x -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_many.log
cat gpurun_out/pytest_many.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --extras or_many_zipf,or_many_sharded > gpurun_out/bench_many.json 2> gpurun_out/bench_many.err
grep -E "or_many|parity" gpurun_out/bench_many.err | tail -12
timeout 200 python tools/scale_probe.py > gpurun_out/scale_product.jsonl 2> gpurun_out/scale_product.err
RB200_LIB=$PWD/croaring_b200/_probe.so timeout 200 python tools/scale_probe.py --strides 1,8 > gpurun_out/scale_probe.jsonl 2> gpurun_out/scale_probe.err
RB200_LIB=$PWD/croaring_b200/_t1.so timeout 200 python tools/scale_probe.py --strides 1,4,8,16 > gpurun_out/scale_t1.jsonl 2> gpurun_out/scale_t1.err
RB200_ORDER_MIN=4000000000 timeout 200 python tools/scale_probe.py --strides 4,8,16 > gpurun_out/scale_noorder.jsonl 2> gpurun_out/scale_noorder.err
tail -3 gpurun_out/scale_*.err
