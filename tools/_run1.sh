timeout 900 python -m pytest tests/test_gpu_serialize.py tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "frozen or relations or deserialize" 2>&1 | tail -15
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
