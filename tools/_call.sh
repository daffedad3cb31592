python tools/sanitize_run.py 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r02o_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/r02o_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r02o_synccheck.log 2>&1; echo synccheck rc=$?; tail -5 gpurun_out/r02o_synccheck.log
