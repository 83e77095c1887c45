#!/bin/bash
# Round-2 GPU call T (final, 1 GPU): full suite, smoke, per-op times, bench (both arms), launch lists,
# DRAM traffic of the dominant kernel, full ncu captures, memcheck of smoke().
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
echo "host cores: $(nproc)  mem: $(free -g | awk '/Mem/{print $2}') GiB" >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 300 python tools/time_ops.py --tag product_final > gpurun_out/ops_final.json 2> gpurun_out/ops_final.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 14 gpurun_out/bench.err
# launch list of the bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/launches_bench.out 2>&1
# DRAM traffic of the dominant kernel over one bench step (9 launches after 3 warm-up steps)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:k_compute_items -s 27 -c 9 --csv --log-file gpurun_out/traffic_compute.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/traffic_compute.out 2>&1
# full captures of the hot kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compute_items -s 1 -c 1 \
   -f -o gpurun_out/prof_compute python tools/profile_target.py pairs 2 > gpurun_out/ncu_compute.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_or_many2 -s 1 -c 1 \
   -f -o gpurun_out/prof_many2_d003 python tools/prof_many.py 0.003 2 > gpurun_out/ncu_many2_d003.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_or_many2 -s 1 -c 1 \
   -f -o gpurun_out/prof_many2_d03 python tools/prof_many.py 0.3 2 > gpurun_out/ncu_many2_d03.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/launches_many2.csv python tools/prof_many.py 0.003 2 > gpurun_out/launches_many2.out 2>&1
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; tail -n 3 gpurun_out/memcheck.log
