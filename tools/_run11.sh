GOGP_PEER_BCAST=1 GOGP_PEER_DEBUG=1 timeout 90 python -u tools/grid_bench.py --size 65536 --gpus 4 --reps 1 > gpurun_out/peer4_dbg4.json 2> gpurun_out/peer4_dbg4.err
echo "rc=$?"; tail -c 700 gpurun_out/peer4_dbg4.json; echo; grep -c "bcast root" gpurun_out/peer4_dbg4.err; grep -i "fail" gpurun_out/peer4_dbg4.err | head -5
for r in 0 1 2 3; do grep "\[peer $r\]" gpurun_out/peer4_dbg4.err | wc -l; grep "\[peer $r\]" gpurun_out/peer4_dbg4.err | tail -2; done
