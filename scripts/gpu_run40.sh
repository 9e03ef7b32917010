python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "readme" 2>&1 | tail -8
