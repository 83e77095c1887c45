#!/bin/bash
# Round-2 GPU call E: per-op timing of the product / norank builds, the whole GPU suite, or_many timings, bench N=1.
mkdir -p gpurun_out
for v in product norank; do
  RB200_LIB=$PWD/croaring_b200/libvar_$v.so timeout 300 python tools/time_ops.py --tag $v > gpurun_out/ops3_$v.json 2> gpurun_out/ops3_$v.err
done
RB200_NO_FUSED=1 RB200_LIB=$PWD/croaring_b200/libvar_product.so timeout 300 python tools/time_ops.py --tag product_nofused --ops and > gpurun_out/ops3_nofused.json 2> gpurun_out/ops3_nofused.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ops3_*.json')):
    try:
        d=json.load(open(f)); o=d['ops']
        print(d['tag'], 'step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:(v['kernel_ms'] if isinstance(v,dict) and 'kernel_ms' in v else v) for k,v in o.items() if 'weather' in k or 'census1881/or' in k or 'dropin' in k or 'successive' in k})
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for d in 0.3 0.03 0.003; do timeout 600 python tools/prof_many.py $d 3 2>&1 | tail -1; done > gpurun_out/many_e.log 2>&1
RB200_OR_MANY_TMA=0 timeout 600 python tools/prof_many.py 0.3 3 2>&1 | tail -1 >> gpurun_out/many_e.log
timeout 900 python tools/prof_many.py 0 3 1000 2>&1 | tail -1 >> gpurun_out/many_e.log
cat gpurun_out/many_e.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -12 gpurun_out/bench.err
