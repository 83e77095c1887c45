#!/bin/bash
# Round-2 GPU call F: product timing after the reverts, or_many2 with the fold prepass (direct / TMA), full suite.
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py --tag product_f > gpurun_out/ops4_product.json 2> gpurun_out/ops4_product.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops4_product.json')); o=d['ops']
print(d['tag'], 'step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:(v['kernel_ms'] if isinstance(v,dict) and 'kernel_ms' in v else v) for k,v in o.items()})
PY
for d in 0.3 0.03 0.003; do timeout 600 python tools/prof_many.py $d 3 2>&1 | tail -1; done > gpurun_out/many_f.log 2>&1
RB200_OR_MANY_TMA=0 timeout 600 python tools/prof_many.py 0.3 3 2>&1 | tail -1 >> gpurun_out/many_f.log
timeout 900 python tools/prof_many.py 0 3 1000 2>&1 | tail -1 >> gpurun_out/many_f.log
cat gpurun_out/many_f.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
