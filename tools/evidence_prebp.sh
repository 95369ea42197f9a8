set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pre_bp --launch-skip 3 -c 1 -f -o gpurun_out/prof_prebp_r1j python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_prebp_j.log 2>&1
ls -la gpurun_out
