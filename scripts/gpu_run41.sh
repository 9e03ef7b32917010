python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/state_hash.py 2>&1 | tail -2
