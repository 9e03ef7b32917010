python -m pytest tests -m gpu -q -x -k "downsample or trajectory or post_process" 2>&1 | tail -15
