echo "=== default connections"
GOGP_PEER_BCAST=1 timeout 60 python -u tools/grid_bench.py --size 32768 --gpus 4 --reps 8 > gpurun_out/peer4_stress_a.json 2> gpurun_out/peer4_stress_a.err; echo "rc=$?"
tail -c 400 gpurun_out/peer4_stress_a.json; echo
echo "=== 32 connections"
CUDA_DEVICE_MAX_CONNECTIONS=32 GOGP_PEER_BCAST=1 timeout 60 python -u tools/grid_bench.py --size 32768 --gpus 4 --reps 8 > gpurun_out/peer4_stress_b.json 2> gpurun_out/peer4_stress_b.err; echo "rc=$?"
tail -c 400 gpurun_out/peer4_stress_b.json; echo
echo "=== spmd 32 connections"
CUDA_DEVICE_MAX_CONNECTIONS=32 GOGP_PEER_BCAST=1 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29537 tools/grid_bench.py --size 32768 --reps 8 > gpurun_out/peer4_stress_c.json 2> gpurun_out/peer4_stress_c.err; echo "rc=$?"
tail -c 400 gpurun_out/peer4_stress_c.json; echo
