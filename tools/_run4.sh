set -x
python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -25
echo "=== RBF 32768"; python tools/eval_rbf.py 32768
echo "=== C3 32768"; python tools/eval_once.py 32768 1
python tools/grid_bench.py --size 32768 --gpus 1 --check
