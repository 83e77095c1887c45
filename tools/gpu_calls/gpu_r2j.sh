#!/bin/bash
# Round-2 GPU call J (8 GPUs): bench at N = 8 (strong scaling of the headline, key-sharded or_many with NCCL).
mkdir -p gpurun_out
timeout 380 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
grep "\[bench\]" gpurun_out/bench_n8.err | sort | uniq | head -12; tail -5 gpurun_out/bench_n8.err
head -c 600 gpurun_out/bench_n8.json; echo
