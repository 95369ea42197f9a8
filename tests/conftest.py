import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> dict with mat (csc), priors, synd [B, m] uint8, kwargs dict and the recorded reference outputs."""
    from scipy.sparse import csc_matrix
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    m, n = (int(x) for x in z["shape"])
    indptr, indices = z["indptr"], z["indices"]
    mat = csc_matrix((np.ones(len(indices), dtype=np.uint8), indices, indptr), shape=(m, n))
    out = {k: z[k] for k in z.files}
    out["mat"] = mat
    out["synd"] = np.unpackbits(z["synd"], axis=1)[:, :m]
    out["kwargs"] = eval(str(z["kwargs"]), {"__builtins__": {}}, {"dict": dict, "True": True, "False": False, "None": None})
    for k in ("dec", "bp_decoding", "osd0"):
        if k in out:
            out[k] = np.unpackbits(out[k], axis=1)[:, :n]
    return out


GOLDEN_GDG = ["c1_gdg_sim_mt1", "c1_gdg_default_mt1", "c1_gdg_sim_mt0", "c1_gdg_default_mt0",
              "c2_w0_gdg_mt1", "c2_w1_gdg_mt1", "c2_w4_gdg_mt1", "c2_w0_gdg_mt0", "c2_w1_gdg_mt0", "c2_w4_gdg_mt0",
              "c3_w0_gdg_mt1", "c3_w5_gdg_mt1", "c3_w10_gdg_mt1",
              "c5_w0_gdg_mt0", "c5_w4_gdg_mt0", "c5_w0_gdg_mt1", "c5_w4_gdg_mt1", "c4_w7_gdg_mt1",
              "c3_w5_gdg_mt0",            # headline window, single-thread schedule
              "g144_gdg_mt1"]             # un-windowed [[144,12,12]] DEM (936 x 8784): HBM-streamed BP on the GPU
GOLDEN_OSD = ["c1_osdw_osd_00", "c1_osdw_osd_cs10", "c1_osdw_osd_e6", "c2_w0_osdw_cs10", "c2_w1_osdw_cs10",
              "c2_w4_osdw_cs10", "c3_w5_osdw_cs10", "c5_w0_osdw_cs10", "c5_w4_osdw_cs10", "c4_w7_osdw_cs10",
              "g144_osdw_cs10"]           # IBM.ipynb:122-123 (pre 16, post 1000, osd_cs 10) on the un-windowed [[144,12,12]] DEM
GOLDEN_BPGD = ["c1_bpgd", "c3_w5_bpgd"]


GOLDEN_BP4 = ["c1_bp4_osd_00", "c1_bp4_osd_cs8", "c1_bp4_osd_e5"]
GOLDEN_CAMEL = ["c1_bp4_camel", "c1_bp4_camel_tied"]      # camel_decode; the same layout plus min_pm [B]


def load_golden_bp4(name):
    """bp4_osd fixtures: Hx, Hz (csc), px/py/pz, synd_x/z [B, m], kwargs, dec [B, 2n], conv, bp_iteration, lpr_first16."""
    from scipy.sparse import csc_matrix
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {k: z[k] for k in z.files}
    for b in ("hx", "hz"):
        m, n = (int(x) for x in z[b + "_shape"])
        out[b] = csc_matrix((np.ones(len(z[b + "_indices"]), dtype=np.uint8), z[b + "_indices"], z[b + "_indptr"]), shape=(m, n))
    n = out["hx"].shape[1]
    out["synd_x"] = np.unpackbits(z["synd_x"], axis=1)[:, :out["hx"].shape[0]]
    out["synd_z"] = np.unpackbits(z["synd_z"], axis=1)[:, :out["hz"].shape[0]]
    out["dec"] = np.unpackbits(z["dec"], axis=1)[:, :2 * n]
    out["kwargs"] = eval(str(z["kwargs"]), {"__builtins__": {}}, {"dict": dict, "True": True, "False": False, "None": None})
    return out


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle
