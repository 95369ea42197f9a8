set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or simulation_data" 2>&1 | tail -25 > gpurun_out/pytest_g5.log; cat gpurun_out/pytest_g5.log
