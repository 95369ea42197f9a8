set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "latency_configuration or single_shot or edge_cases or sliding_window_driver" 2>&1 | tail -4
python tools/latency_probe.py 2>&1 | tail -4
SWD_NO_LATENCY_MODE=1 python tools/latency_probe.py 2>&1 | tail -4 | head -2
