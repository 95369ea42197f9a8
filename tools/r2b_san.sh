# compute-sanitizer on the second-session build (pre-BP kernel with the static message layout): memcheck + racecheck, small batches
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_check.py 8 > gpurun_out/sanitize_memcheck_r2b.log 2>&1; echo "memcheck rc=$?"; tail -14 gpurun_out/sanitize_memcheck_r2b.log
SANITIZE_G144=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_check.py 4 > gpurun_out/sanitize_racecheck_r2b.log 2>&1; echo "racecheck rc=$?"; tail -14 gpurun_out/sanitize_racecheck_r2b.log
