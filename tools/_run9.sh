nproc
GOGP_PEER_BCAST=1 GOGP_PEER_DEBUG=1 timeout 80 python tools/grid_bench.py --size 8192 --block 1024 --gpus 4 --reps 0 --lml-only > gpurun_out/peer4_dbg.json 2> gpurun_out/peer4_dbg.err
echo "rc=$?"
tail -3 gpurun_out/peer4_dbg.json
grep -c "bcast root" gpurun_out/peer4_dbg.err
for r in 0 1 2 3; do echo "rank $r:"; grep "\[peer $r\]" gpurun_out/peer4_dbg.err | tail -3; done
grep -i "fail\|error" gpurun_out/peer4_dbg.err | head -5
