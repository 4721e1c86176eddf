python -m pytest tests/test_gpu_optimize.py -m gpu -q 2>&1 | tail -3
for t in 36 74 110 148 222 296; do echo "=== small tiles $t"; GOGP_SMALL_TILES=$t python tools/c2_bench.py | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['phases_ms'], 'eval', round(d['eval_device_ms'],3))"; done
for t in 36 148; do echo "=== N=16384 small tiles $t"; GOGP_SMALL_TILES=$t python tools/eval_once.py 16384 1 | tail -1; done
