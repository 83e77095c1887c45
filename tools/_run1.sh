timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_serialize.py tests/test_gpu_lazy.py -x -q -m gpu --timeout 800 -k "deserialize or lazy_fold_device or heap_synthetic" > gpurun_out/memcheck_new.log 2>&1
tail -5 gpurun_out/memcheck_new.log
