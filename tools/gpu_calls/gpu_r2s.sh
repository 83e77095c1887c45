#!/bin/bash
# Round-2 GPU call S: block-aggregated cursor atomics in plan / finalize: parity + per-op times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_allpairs.py tests/test_gpu_inplace.py tests/test_gpu_lazy.py tests/test_gpu_flip.py tests/test_gpu_properties.py tests/test_gpu_serialize.py -m gpu -x -q --timeout 900 2>&1 | tail -4 > gpurun_out/pytest_s.log
cat gpurun_out/pytest_s.log
timeout 300 python tools/time_ops.py --reps 5 --tag product_s > gpurun_out/ops_s.json 2> gpurun_out/ops_s.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops_s.json')); o=d['ops']
print('step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:(v['kernel_ms'],v['op_ms']) if isinstance(v,dict) and 'kernel_ms' in v else v for k,v in o.items()})
PY
