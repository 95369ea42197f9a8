"""Developer A/B check (run under gpurun): time several builds of libswd_b200.so on the bench workload and
verify that they return identical corrections.

  python tools/ab_bench.py [--batch B] [--steps K] lib1.so [lib2.so ...]     (paths relative to the package dir)
Each library runs in its own process (SWD_LIB selects it); the line printed per library holds shots/s, the
per-kernel CUDA-event times of one profiled step and a checksum of all committed corrections.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(batch, steps, env_note):
    import numpy as np
    import torch
    import bench
    from slidingwindowdecoder_b200.sliding_window import SlidingWindowDecoder
    dev = torch.device("cuda", 0)
    bench.select_workload(os.environ.get("AB_WORKLOAD", "c3_gdg"))
    plan = bench.build_plan()
    swd = SlidingWindowDecoder(plan, decoder=bench.WL["decoder"], device=0, streams=int(os.environ.get("AB_STREAMS", "1")), **bench.WL["kw"])
    nsteps = steps + 2
    det_all, obs_all = bench.gpu_sample(swd, batch * nsteps, 1234)
    det_all = det_all.view(nsteps, batch, -1); obs_all = obs_all.view(nsteps, batch, -1)
    out = swd.decode_device(det_all[0].clone(), obs_all[0].clone(), return_corrections=True)
    torch.cuda.synchronize()
    digest = hashlib.sha1(out["total_e_hat"].cpu().numpy().tobytes()).hexdigest()[:16]
    counts0 = out["counts"].cpu().numpy().tolist()
    swd.decode_device(det_all[1].clone(), obs_all[1].clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2, nsteps):
        swd.decode_device(det_all[i].clone(), obs_all[i].clone())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    decs = swd.unique_decoders()
    for d in decs:
        d.reset_counters(); d.set_profiling(True)
    swd.decode_device(det_all[2].clone(), obs_all[2].clone())
    torch.cuda.synchronize()
    kt = {}
    for d in decs:
        for k, (t, ln) in d.kernel_times().items():
            if ln:
                kt[k] = round(kt.get(k, 0.0) + t, 2)
    print(json.dumps({"lib": os.environ.get("SWD_LIB", "default"), "env": env_note, "shots_per_s": round(batch * steps / (ms / 1e3), 1),
                      "ms_per_step": round(ms / steps, 2), "kernel_ms_one_step": kt, "counts_step0": counts0, "sha1_corr_step0": digest}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--env", action="append", default=[], help="NAME=VALUE[,NAME=VALUE] variant to run with every library")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.worker:
        worker(a.batch, a.steps, os.environ.get("AB_ENV_NOTE", ""))
        return
    pkg = os.path.join(ROOT, "slidingwindowdecoder_b200")
    for lib in (a.libs or ["libswd_b200.so"]):
        for ev in (a.env or [""]):
            env = dict(os.environ, SWD_LIB=os.path.join(pkg, lib), AB_ENV_NOTE=ev)
            for kv in filter(None, ev.split(",")):
                k, v = kv.split("=", 1); env[k] = v
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "--batch", str(a.batch), "--steps", str(a.steps)],
                               env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else f"{lib} [{ev}] FAILED rc={r.returncode}\n{r.stderr[-2000:]}\n")
            sys.stdout.flush()


if __name__ == "__main__":
    main()
