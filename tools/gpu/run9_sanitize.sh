set -x
mkdir -p gpurun_out
which compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/sanitize_memcheck_r2.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck_r2.txt; tail -15 gpurun_out/sanitize_memcheck_r2.txt
