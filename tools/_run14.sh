set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --bc-steps 1 > gpurun_out/bench_g8_v2.json 2> gpurun_out/bench_g8_v2.err; echo "rc=$?"; tail -3 gpurun_out/bench_g8_v2.err
CUDA_DEVICE_MAX_CONNECTIONS=32 GOGP_PEER_BCAST=1 timeout 150 python -u tools/grid_bench.py --size 131072 --gpus 8 --reps 1 > gpurun_out/grid_n131072_g8_peer.json 2> gpurun_out/grid_n131072_g8_peer.err; echo "rc=$?"; tail -3 gpurun_out/grid_n131072_g8_peer.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/restarts_bench.py --size 16384 --restarts 64 --alg lbfgs --iters 1000 --threshold 1e-6 > gpurun_out/c4_n16384_g8.json 2> gpurun_out/c4_n16384_g8.err; echo "rc=$?"; tail -3 gpurun_out/c4_n16384_g8.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_g8_v2.json","gpurun_out/grid_n131072_g8_peer.json","gpurun_out/c4_n16384_g8.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if "block_cyclic" in d:
            b=d["block_cyclic"]; print(f, "value", d["value"], "bc eval_s", b["eval_s"], "factor", b["factor_s"], "sweep", b["sweep_s"], "eff", b["strong_scaling_efficiency"], b["phases_ms"])
        elif "phases_ms" in d:
            print(f, d["phases_ms"], "eval", d["eval_ms"], "wall", d["wall_ms"], "peer bytes", d.get("of_which_peer_copy_engine"))
        else:
            print(f, {k:d[k] for k in ("seconds","wall_seconds","evaluations","evals_per_s_total","best_objective","finite_restarts")})
    except Exception as e:
        print(f, "ERR", e)
PY
