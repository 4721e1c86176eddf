set -x
nvidia-smi -L
python -m pytest tests/test_gpu_grid.py -x -q 2>&1 | tail -15
python tools/grid_bench.py --size 32768 --gpus 1 --check > gpurun_out/grid_n32768_g1.json 2> gpurun_out/grid_n32768_g1.err; tail -3 gpurun_out/grid_n32768_g1.err
timeout 600 python tools/grid_bench.py --size 32768 --gpus 2 --check > gpurun_out/grid_n32768_g2.json 2> gpurun_out/grid_n32768_g2.err; tail -3 gpurun_out/grid_n32768_g2.err
timeout 600 python tools/grid_bench.py --size 65536 --gpus 2 > gpurun_out/grid_n65536_g2.json 2> gpurun_out/grid_n65536_g2.err; tail -3 gpurun_out/grid_n65536_g2.err
cat gpurun_out/grid_*.json
