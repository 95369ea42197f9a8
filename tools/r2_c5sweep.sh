# BASELINE configs[4]: SHYPS r=3 memory experiment, sliding-window GDG vs BP+OSD at p = 1e-3 .. 5e-3
mkdir -p gpurun_out; rm -f gpurun_out/c5_sweep.jsonl
for p in 0.001 0.002 0.003 0.004 0.005; do for w in c5_gdg c5_osd; do
python bench.py --workload $w --p $p --batch 16384 --streams 2 --steps 5 --skip-cpu 2>> gpurun_out/c5_sweep.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$w','p':$p,'shots_per_s':d['value'],'e2e':d['e2e']['value'],'shots':d['results']['shots'],'flagged':d['results']['flagged'],'failed':d['results']['failed'],'ler':d['results']['failed']/d['results']['shots'],'gdg_fraction':d['results']['gdg_fraction'],'kernel_ms':d['roofline']['kernel_ms']}))" | tee -a gpurun_out/c5_sweep.jsonl
done; done
tail -3 gpurun_out/c5_sweep.err
