#!/bin/bash
# Round-2 GPU call I: lean pass-through kernel: per-op timing, full suite, bench N=1, launch list.
mkdir -p gpurun_out
timeout 300 python tools/time_ops.py --tag product_i > gpurun_out/ops7_product.json 2> gpurun_out/ops7_product.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops7_product.json')); o=d['ops']
print(d['tag'], 'step_kernel', d['step_kernel_ms'], 'step_op', d['step_op_ms'], {k:(v['kernel_ms'],v['op_ms']) if isinstance(v,dict) and 'kernel_ms' in v else v for k,v in o.items()})
PY
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -12 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/launches_bench.out 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:k_compute_items -s 27 -c 9 --csv --log-file gpurun_out/traffic_compute.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/traffic_compute.out 2>&1
