set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "packed or large_bb or streamed or g144" 2>&1 | tail -25 > gpurun_out/pytest_g4.log; cat gpurun_out/pytest_g4.log
timeout 900 python bench.py --workload g144_osd --batch 8192 --streams 1 --steps 3 --warmup 3 > gpurun_out/bench_g144.json 2> gpurun_out/bench_g144.err; tail -c 3000 gpurun_out/bench_g144.json; tail -5 gpurun_out/bench_g144.err
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 4000 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
