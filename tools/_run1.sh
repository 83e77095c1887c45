timeout 900 python -m pytest tests/test_gpu_serialize.py -x -q -m gpu --timeout 600 2>&1 | tail -25 > gpurun_out/deser.log
cat gpurun_out/deser.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 --deselect tests/test_gpu_serialize.py 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_serialized']['value'], d['successive']['value'])"
