"""Generates tests/golden/*.npz from the REAL reference (gongaa/SlidingWindowDecoder).

Run in the build container only (it needs /root/reference):
    python tests/golden/make_golden.py
It copies the reference to a scratch directory under /tmp, applies the two mechanical patches
needed by Cython 3.3 / numpy 2 (np.int_t -> np.int64_t, long( -> int(), builds the Cython
extensions there with /usr/bin/gcc, imports them and records decode() outputs on seeded inputs.
Nothing of the reference is copied into this repository; only inputs and outputs are stored.
Inputs are produced by this repository's own host code (codes.py / dem.py / windows.py), so each
fixture carries the window matrix (CSC), the priors and the syndromes it was generated with.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SCRATCH = "/tmp/refbuild/ref"
OUT = os.path.dirname(os.path.abspath(__file__))


def build_reference():
    if not os.path.exists(os.path.join(SCRATCH, "src")) or not any(f.endswith(".so") and "osd_window" in f for f in os.listdir(os.path.join(SCRATCH, "src"))):
        os.makedirs(os.path.dirname(SCRATCH), exist_ok=True)
        subprocess.run(f"rm -rf {SCRATCH} && cp -r /root/reference {SCRATCH} && chmod -R u+w {SCRATCH}", shell=True, check=True)
        subprocess.run("rm -f src/bp_guessing_decoder.cpp src/osd_window.cpp src/bp4_osd.cpp src/mod2sparse.c && "
                       "sed -i 's/np\\.int_t/np.int64_t/g' src/*.pyx src/*.pxd && sed -i 's/= long(/= int(/g' src/*.pyx && "
                       "CC=/usr/bin/gcc CXX=/usr/bin/g++ LDSHARED='/usr/bin/g++ -shared' python3 setup.py build_ext --inplace > build.log 2>&1",
                       shell=True, check=True, cwd=SCRATCH)
    sys.path.insert(0, SCRATCH)


def csc_of(mat):
    from scipy.sparse import csc_matrix
    A = csc_matrix(mat); A.sort_indices()
    return A.shape, A.indptr.astype(np.int32), A.indices.astype(np.int32)


def save(name, mat, priors, synd, kwargs, **outs):
    shape, indptr, indices = csc_of(mat)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), shape=np.array(shape), indptr=indptr, indices=indices,
                        priors=np.asarray(priors, dtype=np.float64), synd=np.packbits(np.asarray(synd, dtype=np.uint8), axis=1),
                        kwargs=np.array(repr(kwargs)), **outs)
    print("wrote", name, "shots", len(synd), {k: v.shape for k, v in outs.items()})


def run_gdg(cls, mat, priors, synd, kwargs):
    dec = cls(mat, channel_probs=priors, **kwargs)
    out = np.zeros((len(synd), mat.shape[1]), dtype=np.uint8)
    conv = np.zeros(len(synd), dtype=np.uint8)
    for i, s in enumerate(synd):
        out[i] = dec.decode(s)
        conv[i] = int(dec.converge)
    return np.packbits(out, axis=1), conv


def run_osd(cls, mat, priors, synd, kwargs, nlpr=8):
    dec = cls(mat, channel_probs=priors, **kwargs)
    n = mat.shape[1]
    out = np.zeros((len(synd), n), dtype=np.uint8); bp = np.zeros_like(out); o0 = np.zeros_like(out)
    conv = np.zeros(len(synd), dtype=np.uint8); pm = np.zeros(len(synd)); it = np.zeros(len(synd), dtype=np.int32)
    lpr = np.zeros((len(synd), n, 4))
    for i, s in enumerate(synd):
        out[i] = dec.decode(s); conv[i] = int(dec.converge); pm[i] = dec.min_pm; it[i] = dec.bp_iteration
        bp[i] = dec.bp_decoding
        if not conv[i]:
            o0[i] = dec.osd0_decoding
        lpr[i] = dec.log_prob_ratios
    return dict(dec=np.packbits(out, axis=1), conv=conv, min_pm=pm, bp_iteration=it, bp_decoding=np.packbits(bp, axis=1),
                osd0=np.packbits(o0, axis=1), lpr_first8=lpr[:nlpr])


def shyps_section(bpgdg_decoder, osd_window):
    """C5: SHYPS r=3 memory experiment, 6 rounds, (3,1) windows of 63 x 588 (SHYPS.ipynb cell 1), p = 3e-3."""
    from slidingwindowdecoder_b200.dem import shyps_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    chk, obs, pri = dem_to_check_matrices(detector_error_model(shyps_memory_circuit(3, 0.003, 6)))
    plan = build_windows(chk, obs, pri, h=21, W=3, F=1, method=0)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 2500, np.random.default_rng(5))
    for wi in (0, len(plan.windows) - 1):
        w = plan.windows[wi]
        s = det[:, w.row0:w.row1]
        s = s[np.nonzero(s.any(axis=1))[0][:500]]
        kw = dict(pre_max_iter=8, post_max_iter=100, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=10)
        save(f"c5_w{wi}_osdw_cs10", w.mat, w.prior, s, kw, **run_osd(osd_window, w.mat, w.prior, s, kw))
        kw = dict(max_iter=8, multi_thread=False)
        d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
        save(f"c5_w{wi}_gdg_mt0", w.mat, w.prior, s, kw, dec=d, conv=c)
        kw = dict(max_iter=8, multi_thread=True)
        d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
        save(f"c5_w{wi}_gdg_mt1", w.mat, w.prior, s, kw, dec=d, conv=c)


def c4_section(bpgdg_decoder, osd_window):
    """C4: [[288,12,18]] circuit level p = 0.003, 18 rounds, (4,1): a middle window 576 x 4896 (BASELINE configs[3]),
    BP+OSD-CS10 (the configuration's decoder) and multi-thread GDG."""
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    code, A, B = bb_code(288)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(bb_memory_circuit(code, A, B, 0.003, 18, z_basis=True)))
    plan = build_windows(chk, obs, pri, code.N, W=4, F=1, method=1)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 400, np.random.default_rng(288))
    w = plan.windows[7]
    s = det[:, w.row0:w.row1]
    s = s[np.nonzero(s.any(axis=1))[0][:120]]
    kw = dict(pre_max_iter=8, post_max_iter=200, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=10)
    save("c4_w7_osdw_cs10", w.mat, w.prior, s, kw, **run_osd(osd_window, w.mat, w.prior, s, kw))
    kw = dict(max_iter=8, multi_thread=True)
    d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s[:80], kw)
    save("c4_w7_gdg_mt1", w.mat, w.prior, s[:80], kw, dec=d, conv=c)


def global144_section(bpgdg_decoder, osd_window, bpgd_decoder):
    """The un-windowed [[144,12,12]] DEM of IBM.ipynb cell 2 (936 x 8784, 30672 edges, p = 0.004, 12 rounds): its messages
    exceed one SM's shared memory, so the product decodes it with the HBM-streamed BP kernel.
    osd_window with the notebook's kwargs (IBM.ipynb:122-123: pre 16, post 1000, osd_cs 10) and multi-thread GDG."""
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    code, A, B = bb_code(144)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(bb_memory_circuit(code, A, B, 0.004, 12, z_basis=True)))
    det, ob, _ = sample_dem(chk, obs, pri, 64, np.random.default_rng(936))
    s = det[np.nonzero(det.any(axis=1))[0][:40]]
    kw = dict(pre_max_iter=16, post_max_iter=1000, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=10)
    save("g144_osdw_cs10", chk, pri, s, kw, **run_osd(osd_window, chk, pri, s, kw, nlpr=2))
    kw = dict(max_iter=16, multi_thread=True)
    d, c = run_gdg(bpgdg_decoder, chk, pri, s[:24], kw)
    save("g144_gdg_mt1", chk, pri, s[:24], kw, dec=d, conv=c)


def noosd_section(osd_window):
    """osd_order = -1 ("BP only", osd_window.pyx:86,192): a [[72,12,6]] (3,1) middle window with a short post-BP so that a
    good part of the shots ends without convergence - decode() then returns bp_decoding with converge = 0."""
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    code, A, B = bb_code(72)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(bb_memory_circuit(code, A, B, 0.005, 6, z_basis=True)))
    plan = build_windows(chk, obs, pri, code.N, W=3, F=1, method=1)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 3000, np.random.default_rng(7272))
    w = plan.windows[1]
    s = det[:, w.row0:w.row1]
    s = s[np.nonzero(s.any(axis=1))[0][:500]]
    kw = dict(pre_max_iter=8, post_max_iter=12, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=-1)
    save("c2_w1_osdw_noosd", w.mat, w.prior, s, kw, **run_osd(osd_window, w.mat, w.prior, s, kw))


def c3_extra_section(bpgdg_decoder, bpgd_decoder):
    """More of the headline configuration's middle window (216 x 1728): single-thread GDG schedule and BPGD."""
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    code, A, B = bb_code(144)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(bb_memory_circuit(code, A, B, 0.003, 12, z_basis=True)))
    plan = build_windows(chk, obs, pri, code.N, W=3, F=1, method=1)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 1500, np.random.default_rng(144))
    w = plan.windows[5]
    s = det[:, w.row0:w.row1]
    s = s[np.nonzero(s.any(axis=1))[0][:400]]
    kw = dict(max_iter=8, multi_thread=False)
    d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
    save("c3_w5_gdg_mt0", w.mat, w.prior, s, kw, dec=d, conv=c)
    kw = dict(max_iter=8, ms_scaling_factor=1.0, max_iter_per_step=6, max_step=25, gd_factor=1.0)
    d, c = run_gdg(bpgd_decoder, w.mat, w.prior, s, kw)
    save("c3_w5_bpgd", w.mat, w.prior, s, kw, dec=d, conv=c)


def bp4_section():
    """bp4_osd (src/bp4_osd.pyx): [[72,12,6]] BB code, depolarizing code-capacity noise, quaternary BP + OSD per basis."""
    from src.bp4_osd import bp4_osd
    from slidingwindowdecoder_b200.codes import bb_code
    code, _, _ = bb_code(72)
    Hx, Hz = code.hx, code.hz
    n = code.N
    rng = np.random.default_rng(404)
    p = 0.09
    px, py, pz = (p / 3 * (1 + 0.2 * rng.random(n)) for _ in range(3))
    shots = 600
    r = rng.random((shots, n))
    isx, isy, isz = r < px, (r >= px) & (r < px + py), (r >= px + py) & (r < px + py + pz)
    ex, ez = (isx | isy).astype(np.int64), (isy | isz).astype(np.int64)
    sx, sz = (ez @ Hx.T % 2).astype(np.uint8), (ex @ Hz.T % 2).astype(np.uint8)
    hx_shape, hx_p, hx_i = csc_of(Hx); hz_shape, hz_p, hz_i = csc_of(Hz)
    for meth, order in (("osd_0", 0), ("osd_cs", 8), ("osd_e", 5)):
        kw = dict(max_iter=24, ms_scaling_factor=0.9, osd_method=meth, osd_order=order)
        d = bp4_osd(np.asarray(Hx), np.asarray(Hz), channel_probs_x=px, channel_probs_y=py, channel_probs_z=pz, **kw)
        dec = np.zeros((shots, 2 * n), dtype=np.uint8); conv = np.zeros(shots, dtype=np.uint8); it = np.zeros(shots, dtype=np.int32)
        lpr = np.zeros((shots, n, 3))
        for i in range(shots):
            dec[i] = d.decode(sx[i], sz[i]).reshape(-1); conv[i] = int(d.converge); it[i] = d.bp_iteration; lpr[i] = d.log_prob_ratios
        name = f"c1_bp4_{meth}{order}"
        np.savez_compressed(os.path.join(OUT, name + ".npz"), hx_shape=np.array(hx_shape), hx_indptr=hx_p, hx_indices=hx_i,
                            hz_shape=np.array(hz_shape), hz_indptr=hz_p, hz_indices=hz_i, px=px, py=py, pz=pz,
                            synd_x=np.packbits(sx, axis=1), synd_z=np.packbits(sz, axis=1), kwargs=np.array(repr(kw)),
                            dec=np.packbits(dec, axis=1), conv=conv, bp_iteration=it, lpr_first16=lpr[:16])
        print("wrote", name, "shots", shots, "converged", int(conv.sum()))


def camel_section():
    """bp4_osd.camel_decode (src/bp4_osd.pyx:223-248): (a) the [[72,12,6]] matrices as they are, (b) a CAMEL-shaped pair
    Hx = (H1 | 1), Hz = (H2 | 1) (Misc.ipynb cell 6): an all-ones last column ties every check to the pinned qubit.
    A fresh decoder per shot, so a shot on which none of the four runs converges returns the zero-initialised buffer."""
    from src.bp4_osd import bp4_osd
    from slidingwindowdecoder_b200.codes import bb_code
    code, _, _ = bb_code(72)
    for tag, tied in (("c1_bp4_camel", False), ("c1_bp4_camel_tied", True)):
        Hx, Hz = np.asarray(code.hx).astype(np.int64), np.asarray(code.hz).astype(np.int64)
        if tied:
            Hx = np.hstack([Hx, np.ones((Hx.shape[0], 1), dtype=np.int64)]); Hz = np.hstack([Hz, np.ones((Hz.shape[0], 1), dtype=np.int64)])
        n = Hx.shape[1]
        rng = np.random.default_rng(505 + tied)
        p = 0.08
        px, py, pz = (p / 3 * (1 + 0.2 * rng.random(n)) for _ in range(3))
        shots = 400
        r = rng.random((shots, n))
        isx, isy, isz = r < px, (r >= px) & (r < px + py), (r >= px + py) & (r < px + py + pz)
        ex, ez = (isx | isy).astype(np.int64), (isy | isz).astype(np.int64)
        sx, sz = (ez @ Hx.T % 2).astype(np.uint8), (ex @ Hz.T % 2).astype(np.uint8)
        hx_shape, hx_p, hx_i = csc_of(Hx); hz_shape, hz_p, hz_i = csc_of(Hz)
        kw = dict(max_iter=24, ms_scaling_factor=0.9, osd_method="osd_0", osd_order=0)
        dec = np.zeros((shots, 2 * n), dtype=np.uint8); conv = np.zeros(shots, dtype=np.uint8); it = np.zeros(shots, dtype=np.int32)
        lpr = np.zeros((shots, n, 3)); pm = np.zeros(shots)
        for i in range(shots):
            d = bp4_osd(Hx, Hz, channel_probs_x=px, channel_probs_y=py, channel_probs_z=pz, **kw)
            dec[i] = d.camel_decode(sx[i], sz[i]).reshape(-1); conv[i] = int(d.converge); it[i] = d.bp_iteration
            lpr[i] = d.log_prob_ratios; pm[i] = d.min_pm
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), hx_shape=np.array(hx_shape), hx_indptr=hx_p, hx_indices=hx_i,
                            hz_shape=np.array(hz_shape), hz_indptr=hz_p, hz_indices=hz_i, px=px, py=py, pz=pz,
                            synd_x=np.packbits(sx, axis=1), synd_z=np.packbits(sz, axis=1), kwargs=np.array(repr(kw)),
                            dec=np.packbits(dec, axis=1), conv=conv, bp_iteration=it, lpr_first16=lpr[:16], min_pm=pm)
        print("wrote", tag, "shots", shots, "converged", int(conv.sum()))


def main():
    build_reference()
    if len(sys.argv) > 1 and sys.argv[1] in ("bp4", "camel"):
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        bp4_section() if sys.argv[1] == "bp4" else camel_section()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "c4":
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        from src.bp_guessing_decoder import bpgdg_decoder
        from src.osd_window import osd_window
        c4_section(bpgdg_decoder, osd_window)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "noosd":
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        from src.osd_window import osd_window
        noosd_section(osd_window)
        return
    if len(sys.argv) > 1 and sys.argv[1] in ("g144", "c3x"):
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        from src.bp_guessing_decoder import bpgdg_decoder, bpgd_decoder
        from src.osd_window import osd_window
        global144_section(bpgdg_decoder, osd_window, bpgd_decoder) if sys.argv[1] == "g144" else c3_extra_section(bpgdg_decoder, bpgd_decoder)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "shyps":
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 2)
        from src.bp_guessing_decoder import bpgdg_decoder
        from src.osd_window import osd_window
        shyps_section(bpgdg_decoder, osd_window)
        return
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)      # the reference prints "Error setting thread affinity" per thread on small hosts
    from src.bp_guessing_decoder import bpgdg_decoder, bpgd_decoder
    from src.osd_window import osd_window
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    from slidingwindowdecoder_b200.windows import build_windows
    from slidingwindowdecoder_b200.sliding_window import sample_dem

    # ---- C1: [[72,12,6]] code capacity, p = 0.05 (non-uniform priors so that pm ties cannot occur)
    code, A, B = bb_code(72)
    H = code.hx
    rng = np.random.default_rng(20261017)
    shots = 1500
    err = (rng.random((shots, code.N)) < 0.05).astype(np.int64)
    synd = (err @ H.T % 2).astype(np.uint8)
    priors = 0.05 * (1 + 0.3 * rng.random(code.N))
    sim = dict(max_iter_per_step=6, gdg_factor=0.625, max_step=40, max_tree_depth=4, max_side_depth=20, max_tree_branch_step=30,
               max_side_branch_step=20, low_error_mode=True, max_iter=24, ms_scaling_factor=0.625, new_n=72)
    for mt in (True, False):
        kw = dict(sim, multi_thread=mt)
        d, c = run_gdg(bpgdg_decoder, H, priors, synd, kw)
        save(f"c1_gdg_sim_mt{int(mt)}", H, priors, synd, kw, dec=d, conv=c)
        kw = dict(max_iter=8, multi_thread=mt)
        d, c = run_gdg(bpgdg_decoder, H, priors, synd, kw)
        save(f"c1_gdg_default_mt{int(mt)}", H, priors, synd, kw, dec=d, conv=c)
    kw = dict(max_iter=8, ms_scaling_factor=1.0, max_iter_per_step=6, max_step=25, gd_factor=1.0)
    d, c = run_gdg(bpgd_decoder, H, priors, synd, kw)
    save("c1_bpgd", H, priors, synd, kw, dec=d, conv=c)
    for meth, order in (("osd_0", 0), ("osd_cs", 10), ("osd_e", 6)):
        kw = dict(pre_max_iter=8, post_max_iter=50, ms_scaling_factor=0.8, osd_method=meth, osd_order=order)
        save(f"c1_osdw_{meth}{order}", H, priors, synd, kw, **run_osd(osd_window, H, priors, synd, kw))
    # uniform priors (the published configuration): converge flag only is exact (ties), keep for statistics
    pri_u = np.ones(code.N) * 0.05
    kw = dict(sim, multi_thread=False)
    d, c = run_gdg(bpgdg_decoder, H, pri_u, synd, kw)
    save("c1_gdg_sim_uniform_mt0", H, pri_u, synd, kw, dec=d, conv=c)

    # ---- C2: [[72,12,6]] circuit level p = 0.003, 6 rounds, (3,1): windows 0 (first), 1 (middle), 4 (last)
    circ = bb_memory_circuit(code, A, B, 0.003, 6, z_basis=True)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(circ))
    plan = build_windows(chk, obs, pri, code.N, W=3, F=1, method=1)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 3000, np.random.default_rng(72))
    for wi in (0, 1, len(plan.windows) - 1):
        w = plan.windows[wi]
        s = det[:, w.row0:w.row1]
        keep = np.nonzero(s.any(axis=1))[0][:600]         # non-trivial syndromes
        s = s[keep]
        kw = dict(pre_max_iter=8, post_max_iter=200, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=10)
        save(f"c2_w{wi}_osdw_cs10", w.mat, w.prior, s, kw, **run_osd(osd_window, w.mat, w.prior, s, kw))
        kw = dict(max_iter=8, multi_thread=True)
        d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
        save(f"c2_w{wi}_gdg_mt1", w.mat, w.prior, s, kw, dec=d, conv=c)
        kw = dict(max_iter=8, multi_thread=False)
        d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
        save(f"c2_w{wi}_gdg_mt0", w.mat, w.prior, s, kw, dec=d, conv=c)

    # ---- C3: [[144,12,12]] circuit level p = 0.003, 12 rounds, (3,1): windows 0 and 5, GDG
    code, A, B = bb_code(144)
    circ = bb_memory_circuit(code, A, B, 0.003, 12, z_basis=True)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(circ))
    plan = build_windows(chk, obs, pri, code.N, W=3, F=1, method=1)
    det, ob, _ = sample_dem(plan.chk, plan.obs, plan.priors, 1500, np.random.default_rng(144))
    for wi in (0, 5, len(plan.windows) - 1):
        w = plan.windows[wi]
        s = det[:, w.row0:w.row1]
        keep = np.nonzero(s.any(axis=1))[0][:400]
        s = s[keep]
        kw = dict(max_iter=8, multi_thread=True)
        d, c = run_gdg(bpgdg_decoder, w.mat, w.prior, s, kw)
        save(f"c3_w{wi}_gdg_mt1", w.mat, w.prior, s, kw, dec=d, conv=c)
    w = plan.windows[5]
    s = det[:, w.row0:w.row1]; s = s[np.nonzero(s.any(axis=1))[0][:200]]
    kw = dict(pre_max_iter=8, post_max_iter=100, ms_scaling_factor=1.0, osd_method="osd_cs", osd_order=10)
    save("c3_w5_osdw_cs10", w.mat, w.prior, s, kw, **run_osd(osd_window, w.mat, w.prior, s, kw))
    shyps_section(bpgdg_decoder, osd_window)
    c4_section(bpgdg_decoder, osd_window)
    c3_extra_section(bpgdg_decoder, bpgd_decoder)
    global144_section(bpgdg_decoder, osd_window, bpgd_decoder)
    noosd_section(osd_window)
    bp4_section()
    camel_section()


if __name__ == "__main__":
    main()
