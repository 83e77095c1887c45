#!/bin/bash
# Round-2 GPU call O: or_many2 direct path variants (4 consecutive vectors per thread, sparse atomics).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_many_index.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 900 -k "or_many" 2>&1 | tail -4 > gpurun_out/pytest_many.log
cat gpurun_out/pytest_many.log
for d in 0.003 0.01 0.03 0.1; do
  for lib in libroaring_b200 _m3 _f64 _f512; do
    echo "== $lib d=$d" >> gpurun_out/many_variants.log
    RB200_LIB=$PWD/croaring_b200/$lib.so timeout 300 python tools/prof_many.py $d 4 2>&1 | tail -n 2 >> gpurun_out/many_variants.log
  done
done
cat gpurun_out/many_variants.log
