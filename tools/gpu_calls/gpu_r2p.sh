#!/bin/bash
# Round-2 GPU call P: or_many flat-all (merge threshold variant); k_compute_items sparse scatter variants.
mkdir -p gpurun_out
rm -f gpurun_out/many_variants.log
for d in 0.01 0.03 0.1 0.3; do
  for lib in libroaring_b200 _mm1k; do
    echo "== $lib d=$d" >> gpurun_out/many_variants.log
    RB200_LIB=$PWD/croaring_b200/$lib.so timeout 300 python tools/prof_many.py $d 4 2>&1 | tail -n 1 >> gpurun_out/many_variants.log
  done
done
echo "== product direct (RB200_OR_MANY_TMA=0) d=0.3" >> gpurun_out/many_variants.log
RB200_OR_MANY_TMA=0 timeout 300 python tools/prof_many.py 0.3 4 2>&1 | tail -n 1 >> gpurun_out/many_variants.log
echo "== product TMA (RB200_OR_MANY_TMA=1) d=0.1" >> gpurun_out/many_variants.log
RB200_OR_MANY_TMA=1 timeout 300 python tools/prof_many.py 0.1 4 2>&1 | tail -n 1 >> gpurun_out/many_variants.log
cat gpurun_out/many_variants.log
for lib in libroaring_b200 _sp1k _sp4k; do
  RB200_LIB=$PWD/croaring_b200/$lib.so timeout 300 python tools/time_ops.py --ops and,or,xor --reps 5 --tag $lib > gpurun_out/ops_$lib.json 2> gpurun_out/ops_$lib.err
done
python - <<'PY'
import json
for lib in ("libroaring_b200","_sp1k","_sp4k"):
    try:
        d=json.load(open(f'gpurun_out/ops_{lib}.json')); o=d['ops']
        print(lib, 'step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:v['kernel_ms'] for k,v in o.items() if isinstance(v,dict) and 'kernel_ms' in v})
    except Exception as e: print(lib, 'failed', e)
PY
