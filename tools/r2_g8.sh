set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --total-shots 2000000 > gpurun_out/bench_strong_n1.json 2> gpurun_out/bench_strong.err; cat gpurun_out/bench_strong_n1.json; tail -3 gpurun_out/bench_strong.err
bash tools/bench_other.sh 2>&1 | tail -8
