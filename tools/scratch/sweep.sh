for b in 16384 32768 65536; do for s in 2 3 4; do
echo "== batch $b streams $s"; timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --batch $b --streams $s 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])
"; done; done
