# bench lines of the other BASELINE configurations (parity-test cases, not the headline): same flags as the lines kept in profiles/
mkdir -p gpurun_out
python bench.py --workload c1_gdg --batch 65536 --streams 2 --skip-cpu > gpurun_out/bench_c1_gdg.json 2> gpurun_out/bench_c1.err
python bench.py --workload c2_osd --batch 16384 --streams 2 --skip-cpu > gpurun_out/bench_c2_osd.json 2> gpurun_out/bench_c2.err
python bench.py --workload c4_osd --batch 4096 --streams 2 --steps 3 --skip-cpu > gpurun_out/bench_c4_osd.json 2> gpurun_out/bench_c4.err
python bench.py --workload c5_gdg --batch 16384 --streams 2 --skip-cpu > gpurun_out/bench_c5_gdg.json 2> gpurun_out/bench_c5.err
python bench.py --workload c5_osd --batch 16384 --streams 2 --skip-cpu > gpurun_out/bench_c5_osd.json 2> gpurun_out/bench_c5o.err
for f in c1_gdg c2_osd c4_osd c5_gdg c5_osd; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/bench_{f}.json') if l.startswith('{')][-1]); print(f, d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])
except Exception as e: print(f, 'FAILED', e)
PY
done
