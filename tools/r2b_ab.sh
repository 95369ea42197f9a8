# A/B of two library builds on C3 / C2 / C4 (run under gpurun): bash tools/r2b_ab.sh libA.so libB.so
set -x
mkdir -p gpurun_out
python tools/ab_bench.py --batch 16384 --steps 3 "$@" | tee gpurun_out/ab_c3.jsonl
AB_WORKLOAD=c2_osd python tools/ab_bench.py --batch 32768 --steps 3 "$@" | tee gpurun_out/ab_c2.jsonl
AB_WORKLOAD=c4_osd python tools/ab_bench.py --batch 4096 --steps 2 "$@" | tee gpurun_out/ab_c4.jsonl
