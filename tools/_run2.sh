set -x
python -m pytest tests/test_gpu_grid.py -x -q 2>&1 | tail -5
timeout 900 python tools/grid_bench.py --size 65536 --gpus 2 > gpurun_out/grid_n65536_g2_v2.json 2> gpurun_out/grid_n65536_g2_v2.err; tail -3 gpurun_out/grid_n65536_g2_v2.err
timeout 900 python tools/grid_bench.py --size 32768 --gpus 1 > gpurun_out/grid_n32768_g1_v2.json 2> gpurun_out/grid_n32768_g1_v2.err; tail -3 gpurun_out/grid_n32768_g1_v2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --bc-n 65536 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; tail -5 gpurun_out/bench_g2.err
cat gpurun_out/grid_n65536_g2_v2.json gpurun_out/grid_n32768_g1_v2.json gpurun_out/bench_g2.json
