# round-2 (second session) evidence, run under gpurun: ncu --set full of the pre-BP kernel with the static message layout,
# the same capture with SWD_PRE_NO_LAYOUT=1 (plain CSR order) for the before / after of the bank conflicts, the default
# bench line, and the C2 / C4 lines WITH their cpu_baseline
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pre_bp_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_prebp_r2b python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_prebp_r2b.log 2>&1
SWD_PRE_NO_LAYOUT=1 ncu --set full --clock-control none --import-source on -k regex:pre_bp_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_prebp_r2b_csr python bench.py --steps 1 --warmup 3 --skip-cpu --streams 1 > gpurun_out/ncu_prebp_r2b_csr.log 2>&1
python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -c 300 gpurun_out/bench_r2b.json
python bench.py --workload c2_osd --batch 16384 --streams 2 > gpurun_out/bench_c2_osd_r2b.json 2> gpurun_out/bench_c2_r2b.err
python bench.py --workload c4_osd --batch 4096 --streams 2 --steps 3 > gpurun_out/bench_c4_osd_r2b.json 2> gpurun_out/bench_c4_r2b.err
for f in r2b c2_osd_r2b c4_osd_r2b; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_{f}.json') if l.startswith('{')][-1]); print(f, d['value'], d['e2e']['value'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['kernel_ms'])
except Exception as e: print(f, 'FAILED', e)
PY
done
ls -la gpurun_out | tail -8
