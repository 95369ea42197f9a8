set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_check.py 16 > gpurun_out/sanitize_memcheck_r2.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/sanitize_memcheck_r2.log
SANITIZE_G144=0 timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_check.py 6 > gpurun_out/sanitize_racecheck_r2.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/sanitize_racecheck_r2.log
