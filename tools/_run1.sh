timeout 900 python -m pytest tests/test_gpu_foreach.py tests/test_gpu_bind_host.py -x -q -m gpu --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2> gpurun_out/bench_async.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], d['e2e']['checksum_sum_card'], d['checksum_sum_card'], d['e2e_serialized']['value'])"
tail -3 gpurun_out/bench_async.err
