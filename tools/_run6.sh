set -x
for c in 0 1 2 4 8; do
  echo "=== max ctas $c"
  GOGP_NCCL_MAX_CTAS=$c timeout 600 python tools/grid_bench.py --size 65536 --gpus 4 --reps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['grid'], d['phases_ms'], 'eval', d['eval_ms'], 'chol/gpu', d['cholesky_tflops_per_gpu'], 'sweep/gpu', d['sweep_tflops_per_gpu'], 'comm', d['comm_ms_on_priority_stream'])"
done
