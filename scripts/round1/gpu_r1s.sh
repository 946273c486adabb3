timeout 80 python -m pytest tests/test_driver_gpu.py -x -q 2>&1 | tail -25
