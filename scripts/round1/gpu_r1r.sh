timeout 100 python -m pytest tests/test_fullsize_properties_gpu.py -x -q 2>&1 | tail -15
