#!/bin/bash
# Round-2 GPU call D: class-ordered tickets x kernel variants (A/B timing), then the whole GPU suite.
mkdir -p gpurun_out
for v in product norank merge0; do
  RB200_LIB=$PWD/croaring_b200/libvar_$v.so timeout 300 python tools/time_ops.py --tag $v > gpurun_out/ops2_$v.json 2> gpurun_out/ops2_$v.err
  RB200_ORDER_MIN=99999999999 RB200_LIB=$PWD/croaring_b200/libvar_$v.so timeout 300 python tools/time_ops.py --tag ${v}_noorder > gpurun_out/ops2_${v}_noorder.json 2> gpurun_out/ops2_${v}_noorder.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ops2_*.json')):
    try:
        d=json.load(open(f)); o=d['ops']
        print(d['tag'], 'step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:v['kernel_ms'] for k,v in o.items() if 'kernel_ms' in v and ('weather' in k or 'census1881/or' in k)})
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compute_items -s 1 -c 1 \
   -f -o gpurun_out/prof_compute_r2b python tools/profile_target.py pairs 2 > gpurun_out/ncu_compute_r2b.out 2>&1
for d in 0.3 0.03 0.003; do timeout 600 python tools/prof_many.py $d 3 2>&1 | tail -2; done > gpurun_out/many_tma.log 2>&1
timeout 900 python tools/prof_many.py 0 3 1000 2>&1 | tail -2 >> gpurun_out/many_tma.log
cat gpurun_out/many_tma.log
