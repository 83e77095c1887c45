#!/bin/bash
# Round-2 GPU call V: 2-D (window, chunk) index build: parity + config 5 / config 3 timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_many_index.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 900 2>&1 | tail -4 > gpurun_out/pytest_v.log
cat gpurun_out/pytest_v.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --extras or_many_zipf,or_many_sharded > gpurun_out/bench_many.json 2> gpurun_out/bench_many.err
grep -E "or_many|parity" gpurun_out/bench_many.err | tail -n 9
