set -x
python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -6
timeout 300 python tools/restarts_bench.py --size 16384 --restarts 3 --alg lbfgs --iters 1000 --threshold 1e-6 > gpurun_out/c4_probe.json 2> gpurun_out/c4_probe.err; tail -2 gpurun_out/c4_probe.err; cat gpurun_out/c4_probe.json
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; tail -3 gpurun_out/bench_g1.err; cat gpurun_out/bench_g1.json
python tools/eval_once.py 16384 1 > gpurun_out/plain_eval.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:grad_trace_fast|cov_tile_fast" -c 2 -o gpurun_out/r2_elem_c3_final python tools/eval_once.py 16384 1 > gpurun_out/ncu_elem.log 2>&1
tail -2 gpurun_out/ncu_elem.log
