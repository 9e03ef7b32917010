python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "slab_decomposition" 2>&1 | tail -5
