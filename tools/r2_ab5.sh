set -x
mkdir -p gpurun_out
python tools/ab_bench.py --batch 16384 --steps 3 libswd_noclaim.so libswd_b200.so > gpurun_out/ab5.jsonl 2>&1; cat gpurun_out/ab5.jsonl
AB_STREAMS=3 python tools/ab_bench.py --batch 32768 --steps 3 libswd_noclaim.so libswd_b200.so > gpurun_out/ab5_s3.jsonl 2>&1; cat gpurun_out/ab5_s3.jsonl
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "gdg or bpgd or heavy or c4 or edge or single" 2>&1 | tail -4
