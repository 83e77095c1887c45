#!/bin/bash
# Round-2 GPU call U: warp-level full/empty mbarrier pipeline in the TMA path of k_or_many2 (variant build).
mkdir -p gpurun_out
RB200_LIB=$PWD/croaring_b200/_pipe.so RB200_OR_MANY_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_properties.py -m gpu -x -q --timeout 600 -k "or_many" 2>&1 | tail -4 > gpurun_out/pytest_pipe.log
cat gpurun_out/pytest_pipe.log
rm -f gpurun_out/many_variants.log
for d in 0.3 0.1; do
  for lib in libroaring_b200 _pipe; do
    echo "== $lib TMA=1 d=$d" >> gpurun_out/many_variants.log
    RB200_OR_MANY_TMA=1 RB200_LIB=$PWD/croaring_b200/$lib.so timeout 300 python tools/prof_many.py $d 4 2>&1 | tail -n 1 >> gpurun_out/many_variants.log
  done
done
cat gpurun_out/many_variants.log
