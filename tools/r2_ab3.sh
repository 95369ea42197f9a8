set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
SWD_DEBUG=1 python tools/ab_bench.py --batch 16384 --steps 3 libswd_nodiet.so libswd_b200.so > gpurun_out/ab3.jsonl 2>&1; cat gpurun_out/ab3.jsonl
for s in 2 3 4; do AB_STREAMS=$s python tools/ab_bench.py --batch 32768 --steps 3 libswd_b200.so; done > gpurun_out/ab3_s.jsonl 2>&1; cat gpurun_out/ab3_s.jsonl
AB_STREAMS=3 python tools/ab_bench.py --batch 49152 --steps 3 libswd_b200.so > gpurun_out/ab3_b.jsonl 2>&1; cat gpurun_out/ab3_b.jsonl
