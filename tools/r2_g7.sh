set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "streamed or g144 or large_bb or unwindowed" 2>&1 | tail -8 > gpurun_out/pytest_g7.log; cat gpurun_out/pytest_g7.log
export SWD_WS_BYTES=40000000000
timeout 900 python bench.py --workload g144_osd --batch 75776 --streams 1 --steps 2 --warmup 3 --skip-cpu > gpurun_out/bench_g144.json 2> gpurun_out/bench_g144.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g144.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'][:30], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms'])"; tail -3 gpurun_out/bench_g144.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pre_bp_stream_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_stream_r2c python bench.py --workload g144_osd --batch 75776 --streams 1 --steps 1 --warmup 3 --skip-cpu > gpurun_out/ncu_stream.log 2>&1; tail -2 gpurun_out/ncu_stream.log
