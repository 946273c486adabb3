timeout 60 python -m pytest tests/test_parity_gpu.py -x -q -k "minimal or rejected" 2>&1 | tail -25
