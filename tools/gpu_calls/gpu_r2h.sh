#!/bin/bash
# Round-2 GPU call H (2 GPUs): many-way parity subset on one GPU (packed index entries), then bench at N = 2.
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py tests/test_gpu_properties.py tests/test_gpu_lazy.py tests/test_gpu_dropin_c.py -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_h.log
cat gpurun_out/pytest_h.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/time_ops.py --tag product_h --ops or > gpurun_out/ops6_product.json 2> gpurun_out/ops6_product.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops6_product.json')); o=d['ops']
print(d['tag'], {k:v for k,v in o.items() if 'successive' in k or 'dropin' in k})
PY
for d in 0.3 0.003; do CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/prof_many.py $d 3 2>&1 | tail -1; done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -14 gpurun_out/bench_n2.err
head -c 1200 gpurun_out/bench_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
head -c 300 gpurun_out/bench_ref_n2.json; echo
