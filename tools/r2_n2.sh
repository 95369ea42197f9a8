set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --total-shots 2000000 > gpurun_out/bench_strong_n2.json 2> gpurun_out/bench_strong_n2.err; cat gpurun_out/bench_strong_n2.json | cut -c1-700; tail -3 gpurun_out/bench_strong_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --ref-seconds 20 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cut -c1-300 gpurun_out/bench_ref_n2.json
