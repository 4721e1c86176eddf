python tools/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && timeout 330 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_memcheck_default.log python tools/sanitize_small.py > gpurun_out/san_mem_a.out 2>&1; echo "rc=$?"
tail -4 gpurun_out/r2_memcheck_default.log; tail -3 gpurun_out/san_mem_a.out
GOGP_TMA_MIN_K=128 GOGP_SMALL_TILES=0 timeout 330 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_memcheck_tma.log python tools/sanitize_small.py > gpurun_out/san_mem_b.out 2>&1; echo "rc=$?"
tail -4 gpurun_out/r2_memcheck_tma.log; tail -3 gpurun_out/san_mem_b.out
