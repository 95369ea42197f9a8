set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
./tools/onchip_peak > gpurun_out/onchip_peaks.json 2> gpurun_out/onchip_peaks.err; cat gpurun_out/onchip_peaks.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/ab_bench.py --batch 16384 --steps 3 libswd_base.so libswd_nodiet.so libswd_o7.so libswd_o8.so > gpurun_out/ab1.jsonl 2>&1; cat gpurun_out/ab1.jsonl
AB_STREAMS=3 python tools/ab_bench.py --batch 32768 --steps 3 libswd_base.so libswd_o7.so libswd_o8.so > gpurun_out/ab1_s3.jsonl 2>&1; cat gpurun_out/ab1_s3.jsonl
