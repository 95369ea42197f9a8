set -x
mkdir -p gpurun_out
./tools/onchip_peak > gpurun_out/onchip_peaks.json 2> gpurun_out/onchip_peaks.err; cat gpurun_out/onchip_peaks.json
python tools/ab_bench.py --batch 16384 --steps 3 libswd_base.so libswd_nodiet.so libswd_m7.so libswd_norot.so libswd_b200.so > gpurun_out/ab2.jsonl 2>&1; cat gpurun_out/ab2.jsonl
AB_STREAMS=3 python tools/ab_bench.py --batch 32768 --steps 3 libswd_nodiet.so libswd_norot.so libswd_b200.so > gpurun_out/ab2_s3.jsonl 2>&1; cat gpurun_out/ab2_s3.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:path_kernel --launch-skip 10 -c 10 -f -o gpurun_out/prof_path_r2a python bench.py --steps 1 --warmup 3 --batch 32768 --skip-cpu --streams 1 > gpurun_out/ncu_path_r2a.log 2>&1
tail -3 gpurun_out/ncu_path_r2a.log
ls -la gpurun_out | tail -5
