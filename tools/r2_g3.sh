set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "streamed or unwindowed or sort_select or reset_failure or g144" 2>&1 | tail -25 > gpurun_out/pytest_g3.log; cat gpurun_out/pytest_g3.log
