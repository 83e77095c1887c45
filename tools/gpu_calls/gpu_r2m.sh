#!/bin/bash
# Round-2 GPU call M: per-class tickets + heavy-first classes; or_many hybrid flat arrays; forced index tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_many_index.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_properties.py tests/test_gpu_xor_many.py tests/test_gpu_allpairs.py -m gpu -x -q --timeout 900 2>&1 | tail -6 > gpurun_out/pytest_many.log
cat gpurun_out/pytest_many.log
timeout 200 python tools/scale_probe.py --strides 1,2,8 --ops and,or,xor > gpurun_out/scale_product2.jsonl 2> gpurun_out/scale_product2.err
tail -n 3 gpurun_out/scale_product2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --extras or_many_zipf,or_many_sharded > gpurun_out/bench_many.json 2> gpurun_out/bench_many.err
grep -E "or_many|parity|headline" gpurun_out/bench_many.err | tail -12
