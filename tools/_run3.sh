set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== C2 small tiles default"; python tools/c2_bench.py
echo "=== C2 small tiles off"; GOGP_SMALL_TILES=0 python tools/c2_bench.py
echo "=== C2 small tiles 74"; GOGP_SMALL_TILES=74 python tools/c2_bench.py
echo "=== RBF 32768"; python tools/eval_rbf.py 32768
echo "=== RBF 32768 interp"; GOGP_ELEM_FAST=0 python tools/eval_rbf.py 32768
echo "=== C3 32768"; python tools/eval_once.py 32768 1
echo "=== C3 32768 nn8 minb2"; GOGP_TRACE_NN8_MINB2=1 python tools/eval_once.py 32768 1
python tools/eval_once.py 16384 1 > gpurun_out/plain_eval.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:grad_trace_fast|cov_tile_fast" -c 2 -o gpurun_out/r2_elem_c3 python tools/eval_once.py 16384 1 > gpurun_out/ncu_elem.log 2>&1
tail -3 gpurun_out/ncu_elem.log
