set -x
python -m pytest tests -m gpu -q --maxfail=12 -k "not slab_decomposition" 2>&1 | tail -60
