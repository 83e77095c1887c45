#!/bin/bash
# Round-2 GPU call Q: lane-parallel pass-through tickets + one-atomic-per-value scatter: parity, per-op times, scaling.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_allpairs.py tests/test_gpu_inplace.py tests/test_gpu_lazy.py tests/test_gpu_flip.py tests/test_gpu_properties.py -m gpu -x -q --timeout 900 2>&1 | tail -4 > gpurun_out/pytest_q.log
cat gpurun_out/pytest_q.log
timeout 300 python tools/time_ops.py --reps 5 --tag product_q > gpurun_out/ops_q.json 2> gpurun_out/ops_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops_q.json')); o=d['ops']
print('step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:(v['kernel_ms'],v['op_ms']) if isinstance(v,dict) and 'kernel_ms' in v else v for k,v in o.items()})
PY
timeout 200 python tools/scale_probe.py --strides 2,8 --ops and,or,xor > gpurun_out/scale_q.jsonl 2> gpurun_out/scale_q.err
tail -n 2 gpurun_out/scale_q.err
