"""Host-side logic that needs no GPU: DEM builder known answers, window geometry, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dem144():
    from slidingwindowdecoder_b200.codes import bb_code
    from slidingwindowdecoder_b200.dem import bb_memory_circuit, detector_error_model, dem_to_check_matrices
    code, A, B = bb_code(144)
    chk, obs, pri = dem_to_check_matrices(detector_error_model(bb_memory_circuit(code, A, B, 0.003, 4, z_basis=True)))
    return code, chk, obs, pri


def test_dem_known_answers(dem144):
    """Round Analysis.ipynb:15,322 and Sliding Window GDG.ipynb:485 (outputs of real stim)."""
    from slidingwindowdecoder_b200.windows import build_windows
    code, chk, obs, pri = dem144
    assert chk.shape == (360, 3024)
    rw = np.asarray(chk.sum(axis=1)).ravel()
    cw = np.asarray(chk.sum(axis=0)).ravel()
    assert (rw.max(), cw.max(), rw.min(), cw.min()) == (35, 6, 16, 2)
    plan = build_windows(chk, obs, pri, code.N, W=3, F=1, method=1)
    assert plan.anchors == [(0, 0), (72, 648), (144, 1368), (216, 2088), (288, 2808), (360, 3024)]
    assert abs(plan.noisy_prior[0] - 0.027499817877069083) < 1e-16
    assert [w.mat.shape for w in plan.windows] == [(216, 1656), (216, 1728), (216, 1656)]
    assert [w.mat.nnz for w in plan.windows] == [5544, 5976, 5904]
    assert [w.ncommit for w in plan.windows] == [648, 720, 1656]


def test_shyps_dem_known_answers():
    """SHYPS.ipynb:212 (output of real stim): r=3, 4 rounds -> 105 x 833, max (row, col) weight (44, 9), min (28, 2)."""
    from slidingwindowdecoder_b200.dem import shyps_memory_circuit, detector_error_model, dem_to_check_matrices, shyps_code
    from slidingwindowdecoder_b200.windows import build_windows
    cd = shyps_code(3)
    assert not (cd["S_X"] @ cd["S_Z"].T % 2).any() and not (cd["gauge_X"] @ cd["L_Z"].T % 2).any()
    chk, obs, pri = dem_to_check_matrices(detector_error_model(shyps_memory_circuit(3, 0.001, 4)))
    assert chk.shape == (105, 833) and obs.shape[0] == 9
    rw = np.asarray(chk.sum(axis=1)).ravel()
    cw = np.asarray(chk.sum(axis=0)).ravel()
    assert (rw.max(), cw.max(), rw.min(), cw.min()) == (44, 9, 28, 2)
    plan = build_windows(chk, obs, pri, h=21, W=3, F=1, method=0)
    assert [w.mat.shape for w in plan.windows] == [(63, 588), (63, 588), (63, 441)]


@pytest.mark.parametrize("W,F,method", [(3, 1, 1), (5, 2, 1), (4, 1, 2), (3, 1, 0), (2, 2, 1)])
def test_window_commits_partition_the_columns(dem144, W, F, method):
    """Whatever (W, F, method): the committed column ranges tile the DEM columns exactly once and in order, every window
    has W rounds of detectors (the last one reaches the end), identity columns carry the summed prior of what they merge."""
    from slidingwindowdecoder_b200.windows import build_windows
    code, chk, obs, pri = dem144
    plan = build_windows(chk, obs, pri, code.N, W=W, F=F, method=method)
    h = code.N // 2
    pos = 0
    for w in plan.windows:
        assert w.col0 == pos
        pos += w.ncommit
        assert w.row1 - w.row0 == (W * h if not w.last else w.row1 - w.row0)
        if method != 0 and not w.last:
            ident = w.mat[:, w.ncols_dem:].toarray()
            assert np.array_equal(ident[-h:], np.eye(h, dtype=ident.dtype)) and not ident[:-h].any()
            assert np.allclose(w.prior[w.ncols_dem:], plan.noisy_prior)
    assert pos == plan.chk.shape[1] and plan.windows[-1].last and plan.windows[-1].row1 == plan.chk.shape[0]


def test_dem_sampling_consistency(dem144):
    from slidingwindowdecoder_b200.sliding_window import sample_dem
    code, chk, obs, pri = dem144
    det, ob, err = sample_dem(chk, obs, pri, 50, np.random.default_rng(3))
    assert np.array_equal(det, np.asarray((chk @ err.T.astype(np.int64)).T % 2).astype(np.uint8))


def test_philox_known_answers_and_sampler_restatement(dem144):
    """Random123 kat_vectors for philox4x32-10 pin the generator of the device DEM sampler; the numpy restatement
    of the sampler is self-consistent (det = chk . err) and a pure function of (seed, absolute shot index)."""
    from oracle.philox import philox4x32_10, sample_dem, thresholds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        got = philox4x32_10(*[np.uint32(x) for x in ctr], *key)
        assert tuple(int(x) for x in got) == exp
    code, chk, obs, pri = dem144
    det, ob, err = sample_dem(chk, obs, pri, 64, seed=5, shot_offset=100)
    assert np.array_equal(det, np.asarray((chk @ err.T.astype(np.int64)).T % 2).astype(np.uint8))
    det2, ob2, err2 = sample_dem(chk, obs, pri, 32, seed=5, shot_offset=132)
    assert np.array_equal(err[32:], err2) and np.array_equal(det[32:], det2) and np.array_equal(ob[32:], ob2)
    assert thresholds([0.0, 1.0, 0.5])[1] == 0xFFFFFFFF and thresholds([0.0, 1.0, 0.5])[2] == 0x80000000
    # error weight is statistically what the priors say (5 sigma)
    _, _, e = sample_dem(chk, obs, pri, 4000, seed=9)
    mean, sd = pri.sum(), np.sqrt((pri * (1 - pri)).sum() / 4000)
    assert abs(e.sum(axis=1).mean() - mean) < 5 * sd


def test_bb_codes():
    from slidingwindowdecoder_b200.codes import bb_code, gf2_rank
    for N, K in ((72, 12), (144, 12)):
        code, A, B = bb_code(N)
        assert code.N == N and code.K == K
        assert not ((code.hx @ code.lz.T) % 2).any() and not ((code.hz @ code.lx.T) % 2).any()
        assert gf2_rank(np.vstack([code.hz, code.lz])) == code.rank_hz + K


def test_cabi_exports_every_declared_symbol():
    from slidingwindowdecoder_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "swd_b200.h")).read()
    declared = set(re.findall(r"\b(swd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/swd_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert b"sm_100a" in lib.swd_version()
    assert lib.swd_strerror(-2) == b"unsupported configuration"


def test_config_struct_matches_header():
    from slidingwindowdecoder_b200 import _lib
    # field order of swd_config in the header == ctypes mirror
    header = open(os.path.join(ROOT, "include", "swd_b200.h")).read()
    body = header[header.index("typedef struct swd_config {"):header.index("} swd_config;")]
    names = re.findall(r"^\s*(?:int|double)\s+([a-z_]+);", body, flags=re.M)
    assert names == [f for f, _ in _lib.SwdConfig._fields_]
    body = header[header.index("typedef struct swd_counters {"):header.index("} swd_counters;")]
    names = re.findall(r"^\s*uint64_t\s+([a-z_]+);", body, flags=re.M)
    assert names == [f for f, _ in _lib.SwdCounters._fields_]


def test_constructor_argument_errors_need_no_gpu():
    from slidingwindowdecoder_b200 import bpgdg_decoder, osd_window
    H = np.eye(4, dtype=np.uint8)
    with pytest.raises(TypeError):
        bpgdg_decoder([[1, 0], [0, 1]], channel_probs=[0.1, 0.1])
    with pytest.raises(ValueError):
        bpgdg_decoder(H, channel_probs=[0.1] * 3)
    with pytest.raises(ValueError):
        osd_window(H, channel_probs=[0.1] * 4, osd_method="nonsense")


def test_oracle_header_says_test_infrastructure():
    for f in ("swd_oracle.h", "swd_oracle.c", "oracle.py"):
        assert "TEST INFRASTRUCTURE ONLY" in open(os.path.join(ROOT, "oracle", f)).read()
    # the product package never references the oracle
    pkg = os.path.join(ROOT, "slidingwindowdecoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, fn)).read().lower(), fn


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: stdout must carry exactly one
    JSON line with the contract's keys, the arm's own `cpu_baseline` and an `e2e` that repeats the line's value."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-seconds", "9"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "shots/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["steps"] == 2 and d["warmup"] == 1                      # the contract's steps / warm-up are honoured, not clamped
    sys.path.insert(0, root)
    import argparse, bench
    assert d["config"] == bench.config_block(argparse.Namespace(batch=32768, streams=3))     # the product arm's config, verbatim
    assert "modes_shots_per_s" in d["cpu_baseline"] and d["cpu_baseline"]["nproc"] >= 1


@pytest.mark.parametrize("name", ["c1_gdg_sim_mt1", "c3_w5_gdg_mt1", "c5_w0_gdg_mt1"])
def test_pre_bp_message_layout_is_valid_and_nearly_conflict_free(name):
    """The static shared-memory layout swd_create picks for the full-window BP kernel (host code, no CUDA call): every edge
    gets its own slot inside its row's range, the rows of a half-warp of the check pass start on distinct 8-byte banks, and
    the variable pass is left with a small fraction of the extra wavefronts of the plain CSR order."""
    from conftest import load_golden
    from slidingwindowdecoder_b200 import _lib
    lib = _lib.load()
    mat = load_golden(name)["mat"].tocsc()
    mat.sort_indices()
    m, n = mat.shape
    cp, ri = mat.indptr.astype(np.int32), mat.indices.astype(np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    res = {}
    for opt in (0, 1):
        slot, rs, st = np.zeros(mat.nnz, np.int32), np.zeros(m + 1, np.int32), np.zeros(3, np.int64)
        assert lib.swd_pre_bp_layout(m, n, cp.ctypes.data_as(i32p), ri.ctypes.data_as(i32p), opt, slot.ctypes.data, rs.ctypes.data, st.ctypes.data) == 0
        rl = np.bincount(ri, minlength=m)
        assert len(np.unique(slot)) == mat.nnz and slot.max() < st[0]
        assert np.all(slot >= rs[ri]) and np.all(slot < rs[ri] + rl[ri])
        assert np.all(rs[1:] >= rs[:-1] + rl)                          # rows do not overlap
        res[opt] = (slot, rs, st)
    slot0, rs0, st0 = res[0]
    assert st0[0] == mat.nnz and np.array_equal(rs0, np.concatenate([[0], np.cumsum(np.bincount(ri, minlength=m))]))
    slot1, rs1, st1 = res[1]
    for g0 in range(0, m, 16):                                         # check pass: distinct banks inside a half-warp's rows
        b = rs1[g0:min(g0 + 16, m)] & 15
        assert len(np.unique(b)) == len(b)
    assert st1[0] - mat.nnz <= 16 * ((m + 15) // 16)
    assert st1[2] * 5 <= st0[2]                                        # at least 5x fewer conflicting accesses (measured: 50-100x)
    # independent recount of the extra variable-pass wavefronts from the returned slots
    deg = np.diff(cp)
    vord = np.argsort(-deg, kind="stable")
    extra = 0
    for h0 in range(0, n, 16):
        vs = vord[h0:h0 + 16]
        for k in range(int(deg[vs[0]])):
            banks = [int(slot1[cp[v] + k]) & 15 for v in vs if deg[v] > k]
            c = np.bincount(banks, minlength=16)
            extra += int(np.maximum(c - 1, 0).sum())
    assert extra == st1[2]
