#!/bin/bash
# Round-2 GPU call A: tests, smoke, bench (both arms, N=1).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
echo "host cores: $(nproc)  mem: $(free -g | awk '/Mem/{print $2}') GiB" >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1500 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/gpu.txt; cat gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -12 gpurun_out/bench.err
head -c 600 gpurun_out/bench_ref.json; echo; head -c 1500 gpurun_out/bench.json; echo
