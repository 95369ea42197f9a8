for lib in libswd_b200.so libswd_m320.so; do
SWD_LIB=/root/repo/slidingwindowdecoder_b200/$lib python bench.py --workload c4_osd --batch 4096 --streams 2 --steps 3 --skip-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['value'], d['roofline']['kernel_ms'], d['results'])"
done
