# pre-BP physical message layout: GPU tests, then A/B against the plain CSR order (SWD_PRE_NO_LAYOUT=1), identical corrections expected
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/ab_bench.py --batch 16384 --steps 3 --env "" --env SWD_PRE_NO_LAYOUT=1 libswd_b200.so | tee gpurun_out/ab_prelayout.jsonl
AB_WORKLOAD=c2_osd python tools/ab_bench.py --batch 32768 --steps 3 --env "" --env SWD_PRE_NO_LAYOUT=1 libswd_b200.so | tee gpurun_out/ab_prelayout_c2.jsonl
AB_WORKLOAD=c4_osd python tools/ab_bench.py --batch 4096 --steps 2 --env "" --env SWD_PRE_NO_LAYOUT=1 libswd_b200.so | tee gpurun_out/ab_prelayout_c4.jsonl
