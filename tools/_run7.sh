set -x
timeout 600 python -m pytest tests/test_gpu_grid.py -x -q 2>&1 | tail -5
timeout 600 python tools/grid_bench.py --size 65536 --gpus 2 --reps 1 > gpurun_out/grid_n65536_g2_peer.json 2> gpurun_out/grid_n65536_g2_peer.err; tail -3 gpurun_out/grid_n65536_g2_peer.err
GOGP_PEER_BCAST=0 timeout 600 python tools/grid_bench.py --size 65536 --gpus 2 --reps 1 > gpurun_out/grid_n65536_g2_nccl.json 2> gpurun_out/grid_n65536_g2_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/grid_bench.py --size 65536 --reps 1 --check > gpurun_out/grid_n65536_g2_spmd.json 2> gpurun_out/grid_n65536_g2_spmd.err; tail -5 gpurun_out/grid_n65536_g2_spmd.err
for f in gpurun_out/grid_n65536_g2_peer.json gpurun_out/grid_n65536_g2_nccl.json gpurun_out/grid_n65536_g2_spmd.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['mode'], d['phases_ms'], 'eval', round(d['eval_ms'],1), 'comm', d['comm_ms_on_priority_stream'], 'bytes', d.get('collective_bytes_received_rank0'), d.get('of_which_peer_copy_engine'), d.get('lml_rel_diff'), d.get('grad_rel_diff'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
