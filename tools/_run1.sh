timeout 900 python -X faulthandler -m pytest tests/test_gpu_flip.py tests/test_gpu_bind_host.py -x -q -m gpu --timeout 300 2>&1 | tail -6
