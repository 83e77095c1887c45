timeout 900 python -m pytest tests/test_gpu_foreach.py tests/test_gpu_bind_host.py -x -q -m gpu --timeout 600 2>&1 | tail -5
RB200_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu 2> gpurun_out/bench_trace.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['e2e']), d['e2e_serialized']['value'])"
grep "foreach_many" gpurun_out/bench_trace.err | tail -3
