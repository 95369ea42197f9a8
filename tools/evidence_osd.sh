set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:osd_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_osd_c4_r1j python bench.py --workload c4_osd --batch 4096 --steps 1 --warmup 1 --skip-cpu --streams 1 > gpurun_out/ncu_osd_c4_j.log 2>&1
ls -la gpurun_out | tail -3
