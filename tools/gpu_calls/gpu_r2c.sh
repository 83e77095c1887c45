#!/bin/bash
# Round-2 GPU call C: kernel variants of k_compute_items (A/B timing), parity subset, ncu captures.
mkdir -p gpurun_out
for v in product merge0 norank; do
  RB200_LIB=$PWD/croaring_b200/libvar_$v.so timeout 300 python tools/time_ops.py --tag $v > gpurun_out/ops_$v.json 2> gpurun_out/ops_$v.err
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_allpairs.py tests/test_gpu_inplace.py tests/test_gpu_lazy.py -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_c.log
RB200_LIB=$PWD/croaring_b200/libvar_merge0.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_allpairs.py -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_c_merge0.log
# ncu: or_many second generation at d = 0.3 and d = 0.03 (full size), full capture of k_or_many2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_or_many2 -s 1 -c 1 \
   -f -o gpurun_out/prof_many2_d03 python tools/prof_many.py 0.3 2 > gpurun_out/ncu_many2_d03.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_or_many2 -s 1 -c 1 \
   -f -o gpurun_out/prof_many2_d003 python tools/prof_many.py 0.03 2 > gpurun_out/ncu_many2_d003.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
   --log-file gpurun_out/launches_many2.csv python tools/prof_many.py 0.03 2 > gpurun_out/launches_many2.out 2>&1
# full capture of the dominant pairwise kernel (weather all-pairs OR), product build
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compute_items -s 1 -c 1 \
   -f -o gpurun_out/prof_compute_r2 python tools/profile_target.py pairs 2 > gpurun_out/ncu_compute_r2.out 2>&1
cat gpurun_out/ops_*.json; cat gpurun_out/pytest_c.log gpurun_out/pytest_c_merge0.log; tail -3 gpurun_out/ncu_many2_d03.out
