"""numpy restatement of the device DEM sampler (slidingwindowdecoder_b200/csrc/swd_window.cuh: window_sample_kernel).

TEST INFRASTRUCTURE ONLY (imported by tests/).  Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11) pinned by the Random123 known-answer vectors in tests/test_host_logic.py; the sampler itself restates
what the reference draws through stim's CompiledDemSampler (guessing.py:129-130): one Bernoulli per DEM column.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised: counters are uint32 arrays (broadcastable), keys python ints. -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def thresholds(priors):
    x = np.asarray(priors, dtype=np.float64) * 4294967296.0
    return np.where(x >= 4294967295.0, 0xFFFFFFFF, np.floor(x)).astype(np.uint32)


def sample_dem(chk, obs, priors, shots, seed=0, shot_offset=0):
    """-> (det [shots, num_det], obs [shots, num_obs], err [shots, num_col]) uint8, bit-identical to swd_window_sample."""
    num_col = len(priors)
    thr = thresholds(priors)
    blk = np.arange((num_col + 3) // 4, dtype=np.uint64)[None, :]
    shot = (np.arange(shots, dtype=np.uint64) + np.uint64(shot_offset))[:, None]
    r = philox4x32_10(blk, np.uint64(0), shot & MASK, shot >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    draws = np.stack(r, axis=2).reshape(shots, -1)[:, :num_col]
    err = (draws < thr[None, :]).astype(np.uint8)
    e = err.astype(np.float32)
    det = (np.asarray(e @ chk.T.astype(np.float32)) % 2).astype(np.uint8)
    ob = (np.asarray(e @ obs.T.astype(np.float32)) % 2).astype(np.uint8)
    return det, ob, err
