GOGP_PEER_BCAST=1 GOGP_PEER_DEBUG=1 timeout 60 python tools/grid_bench.py --size 8192 --block 1024 --gpus 4 --reps 0 > gpurun_out/peer4_dbg2.json 2> gpurun_out/peer4_dbg2.err
echo "rc=$?"; tail -c 600 gpurun_out/peer4_dbg2.json; echo; grep -c "bcast root" gpurun_out/peer4_dbg2.err; grep -i "fail" gpurun_out/peer4_dbg2.err | head -3
GOGP_PEER_BCAST=1 GOGP_PEER_DEBUG=1 timeout 70 python tools/grid_bench.py --size 32768 --gpus 4 --reps 0 > gpurun_out/peer4_dbg3.json 2> gpurun_out/peer4_dbg3.err
echo "rc=$?"; tail -c 900 gpurun_out/peer4_dbg3.json; echo; grep -c "bcast root" gpurun_out/peer4_dbg3.err; grep -i "fail" gpurun_out/peer4_dbg3.err | head -3
for r in 0 1 2 3; do grep "\[peer $r\]" gpurun_out/peer4_dbg3.err | tail -1; done
