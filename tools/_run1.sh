timeout 900 python -m pytest tests/test_gpu_lazy.py -x -q -m gpu --timeout 600 2>&1 | tail -25 > gpurun_out/lazy.log
cat gpurun_out/lazy.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 --deselect tests/test_gpu_lazy.py 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | head -c 300
