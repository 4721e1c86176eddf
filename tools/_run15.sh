set -x
python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -12
python tools/sanitize_small.py 2>&1 | tail -6
