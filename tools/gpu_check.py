"""Developer check (run under gpurun): CUDA decoders vs the CPU oracle on code-capacity inputs."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slidingwindowdecoder_b200 import bpgdg_decoder, bpgd_decoder, osd_window
from slidingwindowdecoder_b200.codes import bb_code
from oracle.oracle import Oracle


def compare(name, dec, orc_fn, synd, H):
    t0 = time.time()
    corr, conv, pm = dec.decode_batch(synd, return_pm=True)
    t1 = time.time()
    bad = 0
    first = []
    for i in range(synd.shape[0]):
        e, c, p = orc_fn(synd[i])
        ok = np.array_equal(e.astype(np.uint8), corr[i]) and int(c) == int(conv[i])
        if not ok:
            bad += 1
            if len(first) < 5:
                first.append((i, int(c), int(conv[i]), p, pm[i], int(e.sum()), int(corr[i].sum()),
                              bool(((H @ corr[i].astype(np.int64) + synd[i]) % 2).any())))
    print(f"{name}: {bad}/{synd.shape[0]} mismatches, gpu conv {int(conv.sum())}, gpu time {t1-t0:.3f}s", flush=True)
    for f in first:
        print("   shot %d: conv oracle/gpu %d/%d pm %.4f/%.4f wt %d/%d gpu-synd-bad %s" % f)
    print("   counters", dec.counters(), flush=True)
    return bad


def main():
    shots = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    code, _, _ = bb_code(72)
    H = code.hx
    N = code.N
    rng = np.random.default_rng(7)
    p = 0.05
    err = (rng.random((shots, N)) < p).astype(np.int64)
    synd = (err @ H.T % 2).astype(np.uint8)
    priors = p * (1 + 0.3 * rng.random(N))
    orc = Oracle(H, priors)
    total = 0
    kw = dict(max_iter_per_step=6, gdg_factor=0.625, max_step=40, max_tree_depth=4, max_side_depth=20, max_tree_branch_step=30,
              max_side_branch_step=20, low_error_mode=True, max_iter=24, ms_scaling_factor=0.625, new_n=N, multi_thread=True)
    d = bpgdg_decoder(H, channel_probs=priors, **kw)
    total += compare("gdg sim-params", d, lambda s: orc.bpgdg(s, **kw)[:3], synd, H)
    kw2 = dict(max_iter=8, multi_thread=True)
    d = bpgdg_decoder(H, channel_probs=priors, **kw2)
    total += compare("gdg defaults", d, lambda s: orc.bpgdg(s, **kw2)[:3], synd, H)
    kw3 = dict(max_iter=8, multi_thread=True, new_n=50)
    d = bpgdg_decoder(H, channel_probs=priors, **kw3)
    total += compare("gdg defaults new_n=50", d, lambda s: orc.bpgdg(s, **kw3)[:3], synd, H)
    kw4 = dict(max_iter=8, ms_scaling_factor=1.0, max_iter_per_step=6, max_step=25, gd_factor=1.0)
    d = bpgd_decoder(H, channel_probs=priors, **kw4)
    total += compare("bpgd", d, lambda s: orc.bpgd(s, **kw4)[:3], synd, H)
    print("TOTAL MISMATCHES", total)


if __name__ == "__main__":
    main()
