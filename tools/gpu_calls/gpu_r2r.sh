#!/bin/bash
# Round-2 GPU call R: pass-through ticket size sweep.
mkdir -p gpurun_out
for ct in 0 4 8 16 32; do
  RB200_COPY_TICKET=$ct timeout 300 python tools/time_ops.py --ops or,xor --reps 5 --tag ct$ct > gpurun_out/ops_ct$ct.json 2> gpurun_out/ops_ct$ct.err
done
python - <<'PY'
import json
for ct in (0,4,8,16,32):
    d=json.load(open(f'gpurun_out/ops_ct{ct}.json')); o=d['ops']
    print('ct',ct, 'step_kernel', d['step_kernel_ms'], {k:v['kernel_ms'] for k,v in o.items() if isinstance(v,dict) and 'kernel_ms' in v})
PY
