"""Throughput of bp4_osd.decode / camel_decode through the reference-facing batched call (host buffers in, host buffers
out: the copies are inside the timed region) next to the CPU oracle on one host core.  SURVEY 8(f)-4 measurement line.

  python tools/bench_bp4.py [--shots N]        prints one JSON line per method
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shots", type=int, default=65536)
    ap.add_argument("--cpu-shots", type=int, default=1500)
    a = ap.parse_args()
    from conftest import load_golden_bp4
    from slidingwindowdecoder_b200 import bp4_osd
    from oracle import oracle
    for name, method in (("c1_bp4_osd_cs8", "decode"), ("c1_bp4_camel_tied", "camel_decode")):
        g = load_golden_bp4(name)
        hx, hz = np.asarray(g["hx"].todense()).astype(np.int64), np.asarray(g["hz"].todense()).astype(np.int64)
        n = hx.shape[1]
        rng = np.random.default_rng(9)
        r = rng.random((a.shots, n))
        px, py, pz = g["px"], g["py"], g["pz"]
        isx, isy, isz = r < px, (r >= px) & (r < px + py), (r >= px + py) & (r < px + py + pz)
        ex, ez = (isx | isy).astype(np.int64), (isy | isz).astype(np.int64)
        sx, sz = (ez @ hx.T % 2).astype(np.uint8), (ex @ hz.T % 2).astype(np.uint8)
        dec = bp4_osd(g["hx"], g["hz"], channel_probs_x=px, channel_probs_y=py, channel_probs_z=pz, **g["kwargs"])
        fn = dec.decode_batch if method == "decode" else dec.camel_decode_batch
        fn(sx[:1024], sz[:1024]); fn(sx, sz)                                  # warm-up (buffers, clocks)
        t0 = time.perf_counter(); reps = 3
        for _ in range(reps):
            out = fn(sx, sz)
        dt = (time.perf_counter() - t0) / reps
        orc = oracle.Bp4Oracle(g["hx"], g["hz"], px, py, pz)
        kw = g["kwargs"] if method == "decode" else {k: g["kwargs"][k] for k in ("max_iter", "ms_scaling_factor")}
        ofn = orc.decode if method == "decode" else orc.camel_decode
        t0 = time.perf_counter(); same = 0
        for i in range(a.cpu_shots):
            o = ofn(sx[i], sz[i], **kw)
            same += int(np.array_equal(o["dec"].reshape(-1).astype(np.uint8), out["dec"][i].reshape(-1)))
        dc = time.perf_counter() - t0
        print(json.dumps({"method": f"bp4_osd.{method}", "fixture": name, "n": int(n), "kwargs": g["kwargs"], "shots": a.shots,
                          "gpu_shots_per_s_host_to_host": round(a.shots / dt, 1), "converged": float(out["converge"].mean()),
                          "cpu_oracle_shots_per_s_one_core": round(a.cpu_shots / dc, 1),
                          "agree_with_oracle": f"{same}/{a.cpu_shots}"}), flush=True)


if __name__ == "__main__":
    main()
