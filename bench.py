#!/usr/bin/env python
"""Benchmark of the hot path: sliding-window GDG decoding throughput (decoded shots / s).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
  [[144,12,12]] gross code, circuit-level noise p = 0.003, 12 rounds, sliding window (W, F) = (3, 1)
  -> 11 windows of 216 x (1656 | 1728 | 1656), GDG per window (bpgdg_decoder, max_iter=8, multi_thread tree).
A "step" = one pass of the whole window pipeline over one batch of `--batch` synthetic shots per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
Under torchrun (N > 1) every rank decodes its own batch (weak scaling); the only collective is the
all-reduce of the failure counters.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

_STDOUT = sys.stdout
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] - the configuration the metric is quoted on (default)
    "c3_gdg": dict(N=144, p=0.003, rounds=12, W=3, F=1, method=1, decoder="gdg",
                   kw=dict(max_iter=8, max_iter_per_step=6, max_step=25, max_tree_depth=3, max_side_depth=10,
                           max_tree_branch_step=10, max_side_branch_step=10, multi_thread=True, low_error_mode=False),
                   name="[[144,12,12]] circuit-level p=0.003, 12 rounds, sliding window W=3 F=1 (11 windows), GDG per window",
                   decoder_name="bpgdg_decoder(max_iter=8, multi_thread=True, defaults)"),
    # configs[1]: BP + OSD-CS10 per window (osd_window, the reference's own BP+OSD)
    "c2_osd": dict(N=72, p=0.003, rounds=6, W=3, F=1, method=1, decoder="osd",
                   kw=dict(pre_max_iter=8, post_max_iter=200, osd_method="osd_cs", osd_order=10),
                   name="[[72,12,6]] circuit-level p=0.003, 6 rounds, sliding window W=3 F=1 (5 windows), BP+OSD-CS10 per window",
                   decoder_name="osd_window(pre_max_iter=8, post_max_iter=200, osd_cs, order 10)"),
    # configs[3]: large windows (576 x 4896), BP + OSD
    "c4_osd": dict(N=288, p=0.003, rounds=18, W=4, F=1, method=1, decoder="osd",
                   kw=dict(pre_max_iter=8, post_max_iter=200, osd_method="osd_cs", osd_order=10),
                   name="[[288,12,18]] circuit-level p=0.003, 18 rounds, sliding window W=4 F=1 (16 windows), BP+OSD-CS10 per window",
                   decoder_name="osd_window(pre_max_iter=8, post_max_iter=200, osd_cs, order 10)"),
    # configs[0]: code capacity, the reference's own CPU-runnable case (src/simulation.py:10-99 with its GDG kwargs :66-82);
    # one "window" = the whole hx, observables = hz_perp (logical check of simulation.py:90-91)
    "c1_gdg": dict(code="capacity", N=72, p=0.05, decoder="gdg",
                   kw=dict(max_iter_per_step=6, gdg_factor=0.625, max_step=40, max_tree_depth=4, max_side_depth=20,
                           max_tree_branch_step=30, max_side_branch_step=20, multi_thread=True, low_error_mode=True,
                           max_iter=24, ms_scaling_factor=0.625, new_n=72),
                   name="[[72,12,6]] BB code, code-capacity data-qubit noise p=0.05, GDG decode via simulation.py",
                   decoder_name="bpgdg_decoder(simulation.py:66-82 kwargs, multi_thread=True)"),
    # configs[4]: SHYPS r=3 memory experiment (SHYPS.ipynb cell 1: windows without merged identity columns), GDG and BP+OSD
    "c5_gdg": dict(code="shyps", r=3, p=0.003, rounds=6, W=3, F=1, method=0, decoder="gdg",
                   kw=dict(max_iter=8, max_iter_per_step=6, max_step=25, max_tree_depth=3, max_side_depth=10,
                           max_tree_branch_step=10, max_side_branch_step=10, multi_thread=True, low_error_mode=False),
                   name="SHYPS r=3 memory experiment p=0.003, 6 rounds, sliding window W=3 F=1, GDG per window",
                   decoder_name="bpgdg_decoder(max_iter=8, multi_thread=True, defaults)"),
    "c5_osd": dict(code="shyps", r=3, p=0.003, rounds=6, W=3, F=1, method=0, decoder="osd",
                   kw=dict(pre_max_iter=8, post_max_iter=100, osd_method="osd_cs", osd_order=10),
                   name="SHYPS r=3 memory experiment p=0.003, 6 rounds, sliding window W=3 F=1, BP+OSD-CS10 per window",
                   decoder_name="osd_window(pre_max_iter=8, post_max_iter=100, osd_cs, order 10)"),
}
WL = WORKLOADS["c3_gdg"]
METRIC = {"gdg": "decoded shots/sec (sliding-window GDG)", "osd": "decoded shots/sec (sliding-window BP+OSD)"}


def select_workload(key):
    global WL
    WL = WORKLOADS[key]


def build_plan():
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    if WL.get("code") == "capacity":
        from scipy.sparse import csc_matrix
        from slidingwindowdecoder_b200.windows import WindowPlan, Window
        code, _, _ = bb_code(WL["N"])
        hx, lz = csc_matrix(code.hx), csc_matrix(code.hz_perp)
        pri = np.full(code.N, WL["p"])
        win = Window(0, hx, pri, 0, hx.shape[0], 0, code.N, code.N, True)
        return WindowPlan(hx, lz, pri, [(0, 0), (hx.shape[0], code.N)], [win], code.N // 2, 1, 1)
    if WL.get("code") == "shyps":
        from slidingwindowdecoder_b200.dem import shyps_memory_circuit
        r = WL["r"]
        chk, obs, pri = dem_to_check_matrices(detector_error_model(shyps_memory_circuit(r, WL["p"], WL["rounds"])))
        return build_windows(chk, obs, pri, h=r * (2 ** r - 1), W=WL["W"], F=WL["F"], method=WL["method"])
    code, A, B = bb_code(WL["N"])
    circ = bb_memory_circuit(code, A, B, WL["p"], WL["rounds"], z_basis=True)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(circ))
    return build_windows(chk, obs, pri, code.N, W=WL["W"], F=WL["F"], method=WL["method"])


# ------------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_init(plan_blob, use_ref=False, wl_key="c3_gdg"):
    """use_ref: decode with oracle/_ref (the reference's own bpgd.cpp, 15 std::threads per shot) instead of the C port."""
    import ctypes as C
    from oracle.oracle import Oracle, ref_lib, _p
    select_workload(wl_key)
    _CPU["plan"] = plan_blob
    _CPU["orc"] = [Oracle(w.mat, w.prior) for w in plan_blob.windows]
    _CPU["chkT"] = plan_blob.chk.T.tocsr()
    _CPU["obsT"] = plan_blob.obs.T.tocsr()
    _CPU["ref"] = None
    if use_ref:
        lib = ref_lib()
        lib.ref_gdg_create.restype = C.c_void_p
        kw = WL["kw"]
        hs = []
        for o in _CPU["orc"]:
            h = lib.ref_gdg_create(o.m, o.n, _p(o.cp, C.c_int), _p(o.cr, C.c_int), _p(o.llr, C.c_double), kw["max_iter"],
                                   C.c_double(1.0), kw["max_iter_per_step"], kw["max_step"], kw["max_tree_depth"],
                                   kw["max_side_depth"], kw["max_tree_branch_step"], kw["max_side_branch_step"],
                                   C.c_double(1.0), 0, 0)
            hs.append(C.c_void_p(h))
        _CPU["ref"] = (lib, hs)


def _decode_window(i, synd):
    import ctypes as C
    from oracle.oracle import _p
    if _CPU["ref"] is None:
        if WL["decoder"] == "osd":
            dec, conv, _, _ = _CPU["orc"][i].osd_window_batch(synd, **WL["kw"])
        else:
            dec, conv, _, _ = _CPU["orc"][i].bpgdg_batch(synd, **WL["kw"])
        return dec
    lib, hs = _CPU["ref"]
    o = _CPU["orc"][i]
    s = np.ascontiguousarray(synd.astype(np.int8))
    dec = np.zeros((s.shape[0], o.n), dtype=np.int8)
    conv = np.zeros(s.shape[0], dtype=np.int8)
    lib.ref_gdg_decode_batch(hs[i], _p(s, C.c_int8), C.c_longlong(s.shape[0]), _p(dec, C.c_int8), _p(conv, C.c_int8))
    return dec


def _cpu_decode_shard(args):
    """The reference's sliding-window loop (guessing.py:141-227) on a shard of shots with the C oracle."""
    det, obs = args
    plan = _CPU["plan"]
    B = det.shape[0]
    new_det = det.copy()
    total = np.zeros((B, plan.chk.shape[1]), dtype=np.uint8)
    for w in plan.windows:
        dec = _decode_window(w.index, new_det[:, w.row0:w.row1])
        total[:, w.col0:w.col0 + w.ncommit] = dec[:, :w.ncommit]
        upd = np.asarray(total[:, w.col0:w.col0 + w.ncommit].astype(np.float32) @ _CPU["chkT"][w.col0:w.col0 + w.ncommit].astype(np.float32)) % 2
        new_det = (new_det + upd.astype(np.uint8)) % 2
    flagged = new_det.any(axis=1)
    logical = ((obs + np.asarray(total.astype(np.float32) @ _CPU["obsT"].astype(np.float32)) % 2).astype(np.uint8) % 2).any(axis=1)
    return int(flagged.sum()), int(np.logical_or(flagged, logical).sum())


def sample_host(plan, shots, seed):
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    det, obs, _ = sample_dem(plan.chk, plan.obs, plan.priors, shots, np.random.default_rng(seed))
    return det, obs


class CpuArm:
    """Oracle port on all host cores (process-level sharding over shots: the 'all-core' CPU figure of BASELINE.md §3)."""

    def __init__(self, plan, use_ref=False, procs=None):
        import multiprocessing as mp
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.procs = procs or self.cores
        self.plan = plan
        key = next(k for k, v in WORKLOADS.items() if v is WL)
        self.pool = mp.get_context("fork").Pool(self.procs, initializer=_cpu_init, initargs=(plan, use_ref, key))

    def run(self, det, obs):
        B = det.shape[0]
        nshard = min(B, self.procs * 4)
        idx = np.array_split(np.arange(B), nshard)
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_decode_shard, [(det[i], obs[i]) for i in idx if len(i)])
        dt = time.perf_counter() - t0
        return dt, sum(r[0] for r in res), sum(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def calibrate_cpu_sample(arm, plan, target_s):
    det, obs = sample_host(plan, arm.cores * 4, 999)
    dt, _, _ = arm.run(det, obs)
    rate = det.shape[0] / max(dt, 1e-6)
    return int(max(arm.cores * 4, min(20000, rate * target_s)))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm
def gpu_sample(swd, shots, seed):
    """Independent Bernoulli per DEM column drawn on the device (what CompiledDemSampler draws, guessing.py:129-130):
    the repository's own Philox sampler kernel (swd_window_sample), bit-identical to oracle/philox.py."""
    return swd.sample_device(shots, seed=seed)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs CUDA (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from slidingwindowdecoder_b200.sliding_window import SlidingWindowDecoder
    from slidingwindowdecoder_b200.distributed import reduce_counters, max_over_ranks
    plan = build_plan()
    swd = SlidingWindowDecoder(plan, decoder=WL["decoder"], device=local, streams=args.streams, **WL["kw"])
    B, K, W = args.batch, args.steps, args.warmup
    nsteps = K + W
    torch.backends.cuda.matmul.allow_tf32 = False
    det_all, obs_all = gpu_sample(swd, B * nsteps, 1234 + rank)
    det_all = det_all.view(nsteps, B, -1); obs_all = obs_all.view(nsteps, B, -1)
    h_det = torch.empty(det_all.shape, dtype=torch.uint8, pin_memory=True); h_det.copy_(det_all)
    h_obs = torch.empty(obs_all.shape, dtype=torch.uint8, pin_memory=True); h_obs.copy_(obs_all)
    h_counts = torch.zeros((nsteps, 2), dtype=torch.int64, pin_memory=True)
    decs = swd.unique_decoders()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(i):
        det = det_all[i].clone(); obs = obs_all[i].clone()
        return swd.decode_device(det, obs)["counts"]

    def step_single_stream(i):          # kernels launched back to back on one stream: clean per-kernel event times
        det = det_all[i].clone(); obs = obs_all[i].clone()
        return swd.decode_device(det, obs, streams=1)["counts"]

    def step_e2e(i):
        det = h_det[i].to(dev, non_blocking=True); obs = h_obs[i].to(dev, non_blocking=True)
        out = swd.decode_device(det, obs)
        h_counts[i].copy_(out["counts"], non_blocking=True)
        return out["counts"]

    def timed(fn, profile, K=K):
        for i in range(W):
            fn(i)
        for d in decs:
            d.kernel_times() if profile else None
            d.reset_counters()
            d.set_profiling(profile)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        e0.record()
        for i in range(W, W + K):
            counts += fn(i)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1), dev)                # device time, max over ranks
        counts = reduce_counters(counts, dev)                        # the path's only collective: failure counters
        return ms, np.array(counts)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def read_counters():
        out = {}
        for d in decs:
            for k, v in d.counters().items():
                out[k] = out.get(k, 0) + v
        return out

    ms_res, counts_res = timed(step_resident, False)
    ctr_timed = read_counters()
    ms_e2e, counts_e2e = timed(step_e2e, False)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel CUDA-event times + work counters for the roofline: same batches, one stream (with several streams
    # the kernels of different sub-batches overlap and their individual durations are not additive)
    Kp = min(K, 4)
    ms_prof, _ = timed(step_single_stream, True, K=Kp)
    ktimes = {}
    for d in decs:
        for k, (t, ln) in d.kernel_times().items():
            a = ktimes.setdefault(k, [0.0, 0]); a[0] += t; a[1] += ln
        d.set_profiling(False)
    ctr = read_counters()

    # per-window latency (p50) at the throughput batch and at batch 1
    def window_latency(batch, reps):
        lat = []
        for r in range(reps):
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plan.windows]
            det = det_all[r % nsteps][:batch].clone(); obs = obs_all[r % nsteps][:batch].clone()
            swd.decode_device(det, obs, window_events=ev)
            torch.cuda.synchronize()
            lat += [a.elapsed_time(b) for a, b in ev]
        return float(np.median(lat)), float(np.percentile(lat, 99))
    p50_b, p99_b = window_latency(B, 2)
    p50_1, p99_1 = window_latency(1, 20)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_shots = world * B * K
    value = total_shots / (ms_res / 1e3)
    e2e = total_shots / (ms_e2e / 1e3)
    # ---- roofline of the dominant kernel (path_kernel: GDG branch paths, both phases)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    gdg = WL["decoder"] == "gdg"
    dom = ("path_main", "path_side", "path_trunk") if gdg else ("post_bp",)   # osd_window: masked min-sum on the shortened graph
    path_ms = sum(ktimes.get(k, [0, 0])[0] for k in dom)
    path_launches = sum(ktimes.get(k, [0, 0])[1] for k in dom)
    # B_iter = 4 E w + 2 n_a w + (n_a + m_a)/8 with w = 8, summed over executed iterations
    path_bytes = 32.0 * ctr["path_edge_iters"] + 16.0 * ctr["path_vn_iters"] + (ctr["path_vn_iters"] + ctr["path_cn_iters"]) / 8.0
    pre_bytes = 32.0 * ctr["pre_bp_edge_iters"]
    achieved = path_bytes / (path_ms / 1e3) / 1e9 if path_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get("path_kernel" if gdg else "osd_pipeline", {})
        if tj.get("batch") == B:                 # the capture was taken at this batch size
            traffic = int(tj["dram_bytes_per_launch_avg"]); traffic_src = tj.get("source")
    kernel_ms = {k: round(v[0], 3) for k, v in ktimes.items() if v[1]}
    tot_k = sum(kernel_ms.values()) or 1.0
    roofline = {"kernel": "path_kernel (GDG branch paths: shared-prefix nodes + main/tree + side launches)" if gdg else
                          "post_bp_kernel (masked min-sum on the shortened graph; the bit-packed GF(2) osd_kernel is timed separately in kernel_ms)",
                "bound": "hbm", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": round(path_bytes / max(1, path_launches)),
                "avg_launch_ms": round(path_ms / max(1, path_launches), 4), "launches": path_launches,
                "share_of_kernel_time": round(path_ms / tot_k, 4),
                "note": "messages are held in shared memory, so algorithmic message bytes/s is compared with the HBM "
                        "roof only to show on-chip residency; see smem_achieved_gbs for the binding on-chip roof",
                "smem": {"achieved": round(achieved, 1), "peak": round(148 * 128 * 1.965, 1), "unit": "GB/s",
                         "frac": round(achieved / (148 * 128 * 1.965), 4),
                         "note": "same algorithmic bytes against the shared-memory roof (148 SMs x 128 B/clk x 1.965 GHz)"},
                "pre_bp_achieved_gbs": round(pre_bytes / (ktimes.get("pre_bp", [1e-9, 0])[0] / 1e3) / 1e9, 1) if ktimes.get("pre_bp", [0, 0])[0] else None,
                "kernel_ms": kernel_ms, "kernel_ms_steps": Kp, "single_stream_ms_per_step": round(ms_prof / Kp, 3)}
    # ---- CPU baseline on a bounded sample
    if args.skip_cpu or world > 1:        # the CPU baseline is timed on rank 0 of the 1-GPU run only
        cpu_baseline = None
    else:
        arm = CpuArm(plan)
        nsample = calibrate_cpu_sample(arm, plan, 12.0)
        cdet, cobs = sample_host(plan, nsample, 4321)
        dt, cflag, cfail = arm.run(cdet, cobs)
        arm.close()
        cpu_baseline = {"value": round(nsample / dt, 2), "unit": "shots/s", "cores": arm.cores, "kind": "port",
                        "sample": f"{nsample} shots x {len(plan.windows)} windows through the oracle port (C restatement, gcc -O2), "
                                  f"one process per core, {dt:.1f} s; failed {cfail}/{nsample}"}
    launches = ctr_timed["kernel_launches"] + K * args.streams * (2 * len(plan.windows) + 1)
    line = {
        "metric": METRIC[WL["decoder"]], "value": round(value, 1), "unit": "shots/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(ms_res / K, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WL["name"], "shots_per_step_per_gpu": B, "decoder": WL["decoder_name"],
                   "streams": args.streams,
                   "inputs": "DEM samples drawn on the device (Philox, independent Bernoulli per column); distinct batch per step, "
                             f"{(B * (det_all.shape[2] + obs_all.shape[2]) * nsteps) >> 20} MiB of syndromes in total (> L2), no L2 flush"},
        "e2e": {"value": round(e2e, 1), "unit": "shots/s", "h2d_bytes_per_step": int(B * (det_all.shape[2] + obs_all.shape[2])),
                "d2h_bytes_per_step": 16, "ms_per_step": round(ms_e2e / K, 3)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "window_latency_ms": {"p50_at_batch": round(p50_b, 4), "p99_at_batch": round(p99_b, 4), "p50_batch1": round(p50_1, 4), "p99_batch1": round(p99_1, 4)},
        "results": {"shots": int(total_shots), "flagged": int(counts_res[0]), "failed": int(counts_res[1]),
                    "gdg_fraction": round(ctr["gdg_shots"] / max(1, ctr["shots"]), 4)},
        "counters": ctr_timed,
    }
    print(json.dumps(line), file=_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    plan = build_plan()
    # (i) the reference as shipped: ONE process, bpgd.cpp's own 15 std::threads per decode (oracle/_ref), if it was built
    as_shipped = None
    from oracle.oracle import ref_lib
    if ref_lib() is not None and WL["decoder"] == "gdg":
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2); os.dup2(devnull, 2)          # "Error setting thread affinity" spam on hosts with < 15 cores
        try:
            a1 = CpuArm(plan, use_ref=True, procs=1)
            det, obs = sample_host(plan, 48, 7)
            dt, _, _ = a1.run(det, obs)
            a1.close()
            as_shipped = round(48 / dt, 2)
        finally:
            os.dup2(saved, 2)
    # (ii) all-core: one process per core, each running the C port on a shard of the shots
    arm = CpuArm(plan)
    nsample = calibrate_cpu_sample(arm, plan, 8.0)
    K, W = args.steps, args.warmup
    K = max(1, min(K, 6)); W = max(0, min(W, 1))      # bounded: the whole run must end within minutes
    times, fails, n = [], 0, 0
    for i in range(W + K):
        det, obs = sample_host(plan, nsample, 100 + i)
        dt, fl, fa = arm.run(det, obs)
        if i >= W:
            times.append(dt); fails += fa; n += nsample
    arm.close()
    v = n / sum(times)
    line = {"impl": "reference", "metric": METRIC[WL["decoder"]], "value": round(v, 2), "unit": "shots/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": K, "warmup": W, "ms_per_step": round(1e3 * sum(times) / K, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WL["name"], "shots_per_step": nsample},
            "cpu_baseline": {"value": round(v, 2), "unit": "shots/s", "cores": arm.cores, "kind": "port",
                             "sample": f"{nsample} shots x {len(plan.windows)} windows per step, oracle port (C restatement), one process per core",
                             "reference_as_shipped_shots_per_s": as_shipped,
                             "reference_as_shipped_note": "oracle/_ref = the reference's own bpgd.cpp/mod2sparse.c, one process, "
                                                          "15 std::threads per decode, 48 shots" if as_shipped else "oracle/_ref not built"},
            "e2e": {"value": round(v, 2), "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "results": {"shots": n, "failed": fails}}
    print(json.dumps(line), file=_STDOUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_gdg", choices=sorted(WORKLOADS), help="default: BASELINE.json configs[2], the metric's configuration")
    ap.add_argument("--streams", type=int, default=3, help="concurrent sub-batches per GPU (fills kernel tails)")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: do not time the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    # the contract is ONE JSON line on stdout: libraries that write to fd 1 themselves (NCCL prints its version banner
    # there) are sent to stderr, the line itself goes to the saved descriptor
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
