for h in 1 2 3; do echo "=== C4 handles $h"; python tools/restarts_bench.py --size 16384 --restarts 6 --alg adam --iters 5 --handles $h 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('handles_per_gpu','seconds','wall_seconds','evaluations','evals_per_s_total','finite_restarts')})"; done
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g1_v2.json 2> gpurun_out/bench_g1_v2.err; tail -2 gpurun_out/bench_g1_v2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_g1_v2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['phases_ms']); print(d['configs'])"
