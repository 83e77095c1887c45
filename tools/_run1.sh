timeout 900 python -X faulthandler -m pytest tests/test_gpu_r64.py -x -q -m gpu --timeout 300 2>&1 | tail -12
