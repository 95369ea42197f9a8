# round-2 evidence (run under gpurun; every part keeps gpurun_out below the 64 MiB that travel back)
#   bash tools/evidence_r2.sh a   bench lines (product + reference arm), launch list, ncu of pre_bp / sort_reset
#   bash tools/evidence_r2.sh b   ncu --set full of the 10 path_kernel launches of one window
set -x
mkdir -p gpurun_out
if [ "$1" = "a" ]; then
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 400 gpurun_out/bench_r2.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 450 -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pre_bp_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_prebp_r2 python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_prebp_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sort_reset --launch-skip 3 -c 1 -f -o gpurun_out/prof_sort_r2 python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_sort_r2.log 2>&1
else
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 10 -c 10 -f -o gpurun_out/prof_path_r2 python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_path_r2.log 2>&1
fi
ls -la gpurun_out | tail -12
