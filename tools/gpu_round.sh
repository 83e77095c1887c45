#!/bin/bash
# One gpurun call: tests, bench (both arms), extras, ncu launch list + captures (+ sanitizer).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
echo "host cores: $(nproc)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/bench_extras.py card > gpurun_out/extras_card.json 2> gpurun_out/extras_card.err
timeout 900 python tools/bench_extras.py ormany > gpurun_out/extras_ormany.json 2> gpurun_out/extras_ormany.err
timeout 600 python tools/bench_extras.py xormany > gpurun_out/extras_xormany.json 2> gpurun_out/extras_xormany.err
for w in heap lazy deser; do
  timeout 600 python tools/bench_extras.py $w --steps 5 > gpurun_out/extras_$w.json 2> gpurun_out/extras_$w.err
done
timeout 900 python tools/bench_extras.py sharded --bitmaps ${SHARD_BITMAPS:-200} > gpurun_out/extras_sharded_n1.json 2> gpurun_out/extras_sharded_n1.err
for w in pairs card many; do
  timeout 300 python tools/profile_target.py $w 3 > gpurun_out/target_$w.log 2>&1
done
# launch list of the bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.out 2>&1
# DRAM traffic of the dominant kernel over one bench step (9 launches after 3 warm-up steps)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:k_compute_items -s 27 -c 9 --csv --log-file gpurun_out/traffic_compute.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/traffic_compute.out 2>&1
# full captures of the hot kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compute_items -s 1 -c 1 \
   -f -o gpurun_out/prof_compute python tools/profile_target.py pairs 2 > gpurun_out/ncu_compute.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_card_items -s 1 -c 1 \
   -f -o gpurun_out/prof_card python tools/profile_target.py card 2 > gpurun_out/ncu_card.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_or_many -s 1 -c 1 \
   -f -o gpurun_out/prof_many python tools/bench_extras.py ormany --steps 1 --warmup 1 --densities 0.03 --no-cpu > gpurun_out/ncu_many.out 2>&1
if [ -n "$SANITIZE" ]; then
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1
fi
cat gpurun_out/pytest_gpu.log | tail -2; tail -1 gpurun_out/smoke.log; head -c 400 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err; head -c 300 gpurun_out/bench_ref.json; echo
tail -qn 2 gpurun_out/extras_*.err
