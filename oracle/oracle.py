"""ctypes binding of the CPU oracle (oracle/liboracle.so, oracle/_ref/libswd_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(slidingwindowdecoder_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class GdgParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("ms_scaling_factor", C.c_double),
                ("max_iter_per_step", C.c_int), ("max_step", C.c_int),
                ("max_tree_depth", C.c_int), ("max_side_depth", C.c_int),
                ("max_tree_branch_step", C.c_int), ("max_side_branch_step", C.c_int),
                ("gdg_factor", C.c_double), ("new_n", C.c_int),
                ("multi_thread", C.c_int), ("low_error_mode", C.c_int)]


class OsdParams(C.Structure):
    _fields_ = [("pre_max_iter", C.c_int), ("post_max_iter", C.c_int),
                ("ms_scaling_factor", C.c_double), ("new_n", C.c_int),
                ("osd_method", C.c_int), ("osd_order", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("pre_iters", C.c_int64), ("edge_iters", C.c_int64), ("bp_calls", C.c_int64),
                ("paths_run", C.c_int64), ("paths_converged", C.c_int64), ("stage", C.c_int)]


def build(ref=True):
    """(Re)build liboracle.so and, when /root/reference exists, _ref/libswd_ref.so."""
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"] + (["ref"] if ref else []), check=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.orc_gf2_rank.restype = C.c_int
    return _lib


def ref_lib():
    """The reference's own C++ (bpgd.cpp, mod2sparse*.c/cpp) behind ref_shim.cpp; None if not built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libswd_ref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
    return _ref


def set_bp_method(method):
    """Flavour of the full-window BP of every oracle decode that follows: 'minimum_sum' (the reference) or 'product_sum'
    (restated from ldpc's published algorithm; parity unpinned).  Remember to switch back."""
    lib().orc_set_bp_method(1 if str(method).lower() in ("product_sum", "ps", "1") else 0)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def csc_arrays(pcm):
    """(m, n, colptr int32, rows int32 sorted ascending per column) of a dense or scipy matrix."""
    import scipy.sparse as sp
    A = sp.csc_matrix(pcm)
    A.sort_indices()
    A.eliminate_zeros()
    m, n = A.shape
    return m, n, np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32)


def libm_llr(probs):
    """log((1-p)/p) with the C library's log (bp_guessing_decoder.pyx:46) - numpy's differs by ulps."""
    import math
    return np.array([math.log((1.0 - float(p)) / float(p)) for p in probs], dtype=np.float64)


def gdg_params(m, n, max_iter=50, ms_scaling_factor=1.0, max_iter_per_step=6, max_step=25, max_tree_depth=3,
               max_side_depth=10, max_tree_branch_step=10, max_side_branch_step=10, gdg_factor=1.0, new_n=None,
               multi_thread=False, low_error_mode=False, gd_factor=None, **_):
    nn = min(n, 2 * m) if new_n is None else min(int(new_n), n)
    if gd_factor is not None:
        gdg_factor = gd_factor
    return GdgParams(int(max_iter), float(ms_scaling_factor), int(max_iter_per_step), int(max_step),
                     int(max_tree_depth), int(max_side_depth), int(max_tree_branch_step),
                     int(max_side_branch_step), float(gdg_factor), nn, int(bool(multi_thread)),
                     int(bool(low_error_mode)))


_OSD_METHODS = {"osd_0": 0, "0": 0, "osd0": 0, "osd_e": 1, "1": 1, "osde": 1, "exhaustive": 1, "e": 1,
                "osd_cs": 2, "2": 2, "osdcs": 2, "combination_sweep": 2, "cs": 2}


def osd_params(m, n, pre_max_iter=8, post_max_iter=100, ms_scaling_factor=1.0, new_n=None, osd_method="osd_0",
               osd_order=0, **_):
    nn = min(n, 2 * m) if new_n is None else min(int(new_n), n)
    meth = _OSD_METHODS[str(osd_method).lower()]
    if meth == 0:
        osd_order = 0
    return OsdParams(int(pre_max_iter), int(post_max_iter), float(ms_scaling_factor), nn, meth, int(osd_order))


class Oracle:
    """Holds one window PCM + priors; decode functions mirror the reference decoders."""

    def __init__(self, pcm, channel_probs):
        self.m, self.n, self.cp, self.cr = csc_arrays(pcm)
        self.llr = libm_llr(channel_probs)
        self._rank = None

    @property
    def rank(self):
        if self._rank is None:
            self._rank = lib().orc_gf2_rank(self.m, self.n, _p(self.cp, C.c_int), _p(self.cr, C.c_int))
        return self._rank

    def _synd(self, synd):
        s = np.ascontiguousarray(np.asarray(synd).astype(np.int8))
        assert s.shape[-1] == self.m
        return s

    def bp(self, synd, max_iter, alpha=1.0, hist=None):
        s = self._synd(synd)
        if hist is None:
            hist = np.zeros((self.n, 4))
        dec = np.zeros(self.n, dtype=np.int8)
        it = C.c_int(0)
        conv = lib().orc_bp_decode(self.m, self.n, _p(self.cp, C.c_int), _p(self.cr, C.c_int), _p(self.llr, C.c_double),
                                   _p(s, C.c_int8), int(max_iter), C.c_double(alpha), _p(hist, C.c_double),
                                   _p(dec, C.c_int8), C.byref(it))
        return conv, dec, hist, it.value

    def index_sort(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        cols = np.zeros(len(v), dtype=np.int32)
        lib().orc_index_sort(_p(v, C.c_double), len(v), _p(cols, C.c_int))
        return cols

    def bpgdg(self, synd, **kw):
        """-> (dec int8[n], converge, min_pm, Stats)"""
        s = self._synd(synd)
        P = gdg_params(self.m, self.n, **kw)
        dec = np.zeros(self.n, dtype=np.int8)
        pm = C.c_double(0)
        st = Stats()
        fn = lib().orc_bpgd_decode if kw.get("_bpgd") else lib().orc_bpgdg_decode
        conv = fn(self.m, self.n, _p(self.cp, C.c_int), _p(self.cr, C.c_int), _p(self.llr, C.c_double),
                  _p(s, C.c_int8), C.byref(P), _p(dec, C.c_int8), C.byref(pm), C.byref(st))
        return dec, conv, pm.value, st

    def bpgd(self, synd, **kw):
        return self.bpgdg(synd, _bpgd=True, **kw)

    def bpgdg_batch(self, synd, **kw):
        s = self._synd(synd)
        B = s.shape[0]
        P = gdg_params(self.m, self.n, **kw)
        dec = np.zeros((B, self.n), dtype=np.int8)
        conv = np.zeros(B, dtype=np.int8)
        pm = np.zeros(B)
        st = Stats()
        lib().orc_bpgdg_decode_batch(self.m, self.n, _p(self.cp, C.c_int), _p(self.cr, C.c_int),
                                     _p(self.llr, C.c_double), _p(s, C.c_int8), C.c_int64(B), C.byref(P),
                                     _p(dec, C.c_int8), _p(conv, C.c_int8), _p(pm, C.c_double), C.byref(st))
        return dec, conv, pm, st

    def osd_window(self, synd, **kw):
        """-> dict(dec, converge, min_pm, bp_iteration, bp_decoding, osd0_decoding, osdw_decoding, log_prob_ratios)"""
        s = self._synd(synd)
        P = osd_params(self.m, self.n, **kw)
        n = self.n
        dec, bpd, o0, ow = (np.zeros(n, dtype=np.int8) for _ in range(4))
        hist = np.zeros((n, 4))
        pm = C.c_double(0)
        it = C.c_int(0)
        st = Stats()
        conv = lib().orc_osd_window_decode(self.m, n, _p(self.cp, C.c_int), _p(self.cr, C.c_int),
                                           _p(self.llr, C.c_double), _p(s, C.c_int8), C.byref(P), self.rank,
                                           _p(dec, C.c_int8), _p(bpd, C.c_int8), _p(o0, C.c_int8), _p(ow, C.c_int8),
                                           _p(hist, C.c_double), C.byref(pm), C.byref(it), C.byref(st))
        return dict(dec=dec, converge=conv, min_pm=pm.value, bp_iteration=it.value, bp_decoding=bpd,
                    osd0_decoding=o0, osdw_decoding=ow, log_prob_ratios=hist, stats=st)

    def osd_window_batch(self, synd, **kw):
        s = self._synd(synd)
        B = s.shape[0]
        P = osd_params(self.m, self.n, **kw)
        dec = np.zeros((B, self.n), dtype=np.int8)
        conv = np.zeros(B, dtype=np.int8)
        pm = np.zeros(B)
        st = Stats()
        lib().orc_osd_window_decode_batch(self.m, self.n, _p(self.cp, C.c_int), _p(self.cr, C.c_int),
                                          _p(self.llr, C.c_double), _p(s, C.c_int8), C.c_int64(B), C.byref(P),
                                          self.rank, _p(dec, C.c_int8), _p(conv, C.c_int8), _p(pm, C.c_double),
                                          C.byref(st))
        return dec, conv, pm, st


class Bp4Oracle:
    """bp4_osd (bp4_osd.pyx) on the CPU: quaternary BP over (Hx, Hz) + OSD per basis."""

    def __init__(self, Hx, Hz, channel_probs_x, channel_probs_y, channel_probs_z):
        import math
        self.mx, self.n, self.cpx, self.crx = csc_arrays(Hx)
        self.mz, n2, self.cpz, self.crz = csc_arrays(Hz)
        assert n2 == self.n
        px, py, pz = (np.asarray(a, dtype=np.float64) for a in (channel_probs_x, channel_probs_y, channel_probs_z))
        self.llrx = np.array([math.log((1.0 - (px[v] + py[v] + pz[v])) / px[v]) for v in range(self.n)])
        self.llry = np.array([math.log((1.0 - (px[v] + py[v] + pz[v])) / py[v]) for v in range(self.n)])
        self.llrz = np.array([math.log((1.0 - (px[v] + py[v] + pz[v])) / pz[v]) for v in range(self.n)])
        self.prior_x = np.array([math.log((1.0 - (px[v] + py[v])) / (px[v] + py[v])) for v in range(self.n)])
        self.prior_z = np.array([math.log((1.0 - (pz[v] + py[v])) / (pz[v] + py[v])) for v in range(self.n)])
        L = lib()
        self.rank_x = L.orc_gf2_rank(self.mx, self.n, _p(self.cpx, C.c_int), _p(self.crx, C.c_int))
        self.rank_z = L.orc_gf2_rank(self.mz, self.n, _p(self.cpz, C.c_int), _p(self.crz, C.c_int))

    def decode(self, synd_x, synd_z, max_iter=32, ms_scaling_factor=1.0, osd_method="osd_0", osd_order=0):
        """-> dict(dec [2, n], converge, bp_decoding [2, n], osd0 [2, n], log_prob_ratios [n, 3], bp_iteration)"""
        sx = np.ascontiguousarray(np.asarray(synd_x).astype(np.int8)); sz = np.ascontiguousarray(np.asarray(synd_z).astype(np.int8))
        meth = _OSD_METHODS[str(osd_method).lower()]
        n = self.n
        dec, bp, o0 = (np.zeros(2 * n, dtype=np.int8) for _ in range(3))
        lpr = np.zeros((n, 3)); it = C.c_int(0)
        conv = lib().orc_bp4_osd_decode(self.mx, self.mz, n, _p(self.cpx, C.c_int), _p(self.crx, C.c_int), _p(self.cpz, C.c_int),
                                        _p(self.crz, C.c_int), _p(self.llrx, C.c_double), _p(self.llry, C.c_double),
                                        _p(self.llrz, C.c_double), _p(self.prior_x, C.c_double), _p(self.prior_z, C.c_double),
                                        _p(sx, C.c_int8), _p(sz, C.c_int8), int(max_iter), C.c_double(ms_scaling_factor), meth,
                                        0 if meth == 0 else int(osd_order), self.rank_x, self.rank_z, _p(dec, C.c_int8),
                                        _p(bp, C.c_int8), _p(o0, C.c_int8), _p(lpr, C.c_double), C.byref(it))
        return dict(dec=dec.reshape(2, n), converge=int(conv), bp_decoding=bp.reshape(2, n), osd0=o0.reshape(2, n),
                    log_prob_ratios=lpr, bp_iteration=it.value)


    def camel_decode(self, synd_x, synd_z, max_iter=32, ms_scaling_factor=1.0):
        """-> dict(dec [2, n], converge, min_pm, log_prob_ratios [n, 3] (last run), bp_iteration (last run))"""
        sx = np.ascontiguousarray(np.asarray(synd_x).astype(np.int8)); sz = np.ascontiguousarray(np.asarray(synd_z).astype(np.int8))
        n = self.n
        dec = np.zeros(2 * n, dtype=np.int8); lpr = np.zeros((n, 3)); it = C.c_int(0); pm = C.c_double(0)
        conv = lib().orc_bp4_camel_decode(self.mx, self.mz, n, _p(self.cpx, C.c_int), _p(self.crx, C.c_int), _p(self.cpz, C.c_int),
                                          _p(self.crz, C.c_int), _p(self.llrx, C.c_double), _p(self.llry, C.c_double),
                                          _p(self.llrz, C.c_double), _p(sx, C.c_int8), _p(sz, C.c_int8), int(max_iter),
                                          C.c_double(ms_scaling_factor), _p(dec, C.c_int8), _p(lpr, C.c_double), C.byref(pm), C.byref(it))
        return dict(dec=dec.reshape(2, n), converge=int(conv), min_pm=pm.value, log_prob_ratios=lpr, bp_iteration=it.value)


def sliding_window_reference(plan, det, obs, decode_window):
    """The reference's window loop (guessing.py:141-227) in dense numpy, for tests.
    decode_window(window, synd[B, m]) -> (corr[B, n_win], conv[B]).
    Returns dict(flagged[B] bool, failed[B] bool, total_e_hat[B, num_col], window_unconverged list)."""
    chk = np.asarray(plan.chk.todense()).astype(np.int64)
    ob = np.asarray(plan.obs.todense()).astype(np.int64)
    det = np.asarray(det).astype(np.int64)
    obs = np.asarray(obs).astype(np.int64)
    B = det.shape[0]
    total = np.zeros((B, chk.shape[1]), dtype=np.int64)
    new_det = det.copy()
    unconv = []
    for w in plan.windows:
        synd = new_det[:, w.row0:w.row1]
        corr, conv = decode_window(w, synd)
        total[:, w.col0:w.col0 + w.ncommit] = np.asarray(corr)[:, :w.ncommit]
        unconv.append(int(B - np.asarray(conv).astype(np.int64).sum()))
        new_det = (det + total @ chk.T) % 2
    flagged = new_det.any(axis=1)
    logical = ((obs + total @ ob.T) % 2).any(axis=1)
    return dict(flagged=flagged, failed=np.logical_or(flagged, logical), total_e_hat=total, window_unconverged=unconv)
