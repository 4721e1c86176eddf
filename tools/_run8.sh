set -x
timeout 600 python tools/grid_bench.py --size 65536 --gpus 4 --reps 1 > gpurun_out/grid_n65536_g4_peer.json 2> gpurun_out/grid_n65536_g4_peer.err; tail -3 gpurun_out/grid_n65536_g4_peer.err
GOGP_PEER_BCAST=0 timeout 600 python tools/grid_bench.py --size 65536 --gpus 4 --reps 1 > gpurun_out/grid_n65536_g4_nccl.json 2> gpurun_out/grid_n65536_g4_nccl.err
for f in gpurun_out/grid_n65536_g4_peer.json gpurun_out/grid_n65536_g4_nccl.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['grid'], d['phases_ms'], 'eval', round(d['eval_ms'],1), 'wall', round(d['wall_ms'],1), 'comm', d['comm_ms_on_priority_stream'], 'bytes', d.get('collective_bytes_received_rank0'), d.get('of_which_peer_copy_engine'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
