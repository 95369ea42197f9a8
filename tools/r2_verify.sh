# final verification of a build (run under gpurun): smoke, all GPU tests, the default bench line
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_final.json')); print(d['value'], d['e2e']['value'], d['e2e_corrections']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['window_latency_ms'])"
