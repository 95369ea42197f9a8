set -x
mkdir -p gpurun_out
echo skip tests
export SWD_WS_BYTES=40000000000
for lib in libswd_sm3.so libswd_b200.so; do
SWD_LIB=/root/repo/slidingwindowdecoder_b200/$lib timeout 900 python bench.py --workload g144_osd --batch 75776 --streams 1 --steps 2 --warmup 3 --skip-cpu > gpurun_out/bench_g144_$lib.json 2> gpurun_out/bench_g144.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g144_$lib.json')); print('$lib', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernels'].get('pre_bp',d['roofline'])['frac'])"; tail -2 gpurun_out/bench_g144.err
done
