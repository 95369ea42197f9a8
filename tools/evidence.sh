set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err
tail -c 600 gpurun_out/bench_r1j.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 450 -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/bench_under_ncu_j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 10 -c 10 -f -o gpurun_out/prof_path_r1j python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_full_j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sort_reset --launch-skip 3 -c 1 -f -o gpurun_out/prof_sort_r1j python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_sort_j.log 2>&1
ls -la gpurun_out | tail -8
