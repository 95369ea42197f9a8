# final build of round 2 (run under gpurun):  bash tools/evidence_r2c.sh a | b
#   a  smoke, all GPU tests, the default bench line, compute-sanitizer
#   b  ncu --set full of one window's 10 path_kernel launches and of one pre_bp_kernel launch, launch list, bench lines of the other workloads
set -x
mkdir -p gpurun_out
if [ "$1" = "a" ]; then
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c.json')); print(d['value'], d['e2e']['value'], d['e2e_corrections']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['window_latency_ms'], d['roofline']['kernel_ms'])"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_check.py 8 > gpurun_out/sanitize_memcheck_r2c.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitize_memcheck_r2c.log
SANITIZE_G144=0 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_check.py 4 > gpurun_out/sanitize_racecheck_r2c.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitize_racecheck_r2c.log
else
ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 10 -c 10 -f -o gpurun_out/prof_path_r2c python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_path_r2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pre_bp_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_prebp_r2c python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_prebp_r2c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 450 -c 400 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/bench_under_ncu_r2c.log 2>&1
python bench.py --workload c1_gdg --batch 65536 --streams 2 --skip-cpu > gpurun_out/bench_c1_gdg_r2c.json 2> gpurun_out/bench_c1_r2c.err
python bench.py --workload c2_osd --batch 16384 --streams 2 > gpurun_out/bench_c2_osd_r2c.json 2> gpurun_out/bench_c2_r2c.err
python bench.py --workload c4_osd --batch 4096 --streams 2 --steps 3 > gpurun_out/bench_c4_osd_r2c.json 2> gpurun_out/bench_c4_r2c.err
python bench.py --workload c5_gdg --batch 16384 --streams 2 --skip-cpu > gpurun_out/bench_c5_gdg_r2c.json 2> gpurun_out/bench_c5_r2c.err
python bench.py --workload c5_osd --batch 16384 --streams 2 --skip-cpu > gpurun_out/bench_c5_osd_r2c.json 2> gpurun_out/bench_c5o_r2c.err
SWD_WS_BYTES=40000000000 python bench.py --workload g144_osd --batch 75776 --streams 1 --steps 3 --skip-cpu > gpurun_out/bench_g144_osd_r2c.json 2> gpurun_out/bench_g144_r2c.err
for f in c1_gdg c2_osd c4_osd c5_gdg c5_osd g144_osd; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_{f}_r2c.json') if l.startswith('{')][-1]); print(f, d['value'], d['e2e']['value'], (d.get('cpu_baseline') or {}).get('value'), d['roofline']['kernel_ms'])
except Exception as e: print(f, 'FAILED', e)
PY
done
fi
ls -la gpurun_out | tail -6
