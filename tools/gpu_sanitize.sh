timeout 600 python tools/sanitize_small.py > gpurun_out/r02_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/r02_sanitize_plain.log; exit 1; }
tail -3 gpurun_out/r02_sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r02_sanitize_memcheck.log | head -10; tail -3 gpurun_out/r02_sanitize_memcheck.log
