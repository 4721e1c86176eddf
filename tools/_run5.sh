set -x
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --bc-steps 2 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; tail -5 gpurun_out/bench_g8.err
cat gpurun_out/bench_g8.json
timeout 600 python tools/grid_bench.py --size 131072 --gpus 8 --reps 1 > gpurun_out/grid_n131072_g8_threads.json 2> gpurun_out/grid_n131072_g8_threads.err; tail -3 gpurun_out/grid_n131072_g8_threads.err
cat gpurun_out/grid_n131072_g8_threads.json
