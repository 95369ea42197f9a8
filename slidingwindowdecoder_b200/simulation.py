"""Code-capacity driver: the GDG part of the reference's src/simulation.py:10-99 with batched decoding.

`data_qubit_noise_decoding(code, p, num_shots)` samples iid X errors, decodes all syndromes of hx with
`bpgdg_decoder` (same kwargs as simulation.py:66-82) in ONE batched call and counts logical errors against
`hz_perp`.  The `ldpc.BpOsdDecoder` legs of the reference function are third-party and not reproduced.
"""
import time

import numpy as np

from .decoders import bpgdg_decoder


def data_qubit_noise_decoding(code, p, num_shots=1000, max_step=40, max_tree_step=30, max_iter_per_step=6, seed=None,
                              extra_decoders=(), device=0, verbose=True):
    rng = np.random.default_rng(seed)
    err = (rng.random((num_shots, code.N)) < p).astype(np.uint8)             # simulation.py:15
    syndrome = (err.astype(np.int64) @ code.hx.T % 2).astype(np.uint8)       # simulation.py:16
    priors = np.ones(code.N) * p
    results = {}
    decoders = list(extra_decoders)
    decoders.append(("GDG", bpgdg_decoder(                                   # simulation.py:66-82
        code.hx, channel_probs=priors, max_iter_per_step=max_iter_per_step, gdg_factor=0.625, max_step=max_step,
        max_tree_depth=4, max_side_depth=20, max_tree_branch_step=max_tree_step, max_side_branch_step=max_step - 20,
        multi_thread=True, low_error_mode=True, max_iter=24, ms_scaling_factor=0.625, new_n=code.N, device=device)))
    for name, dec in decoders:
        t0 = time.perf_counter()
        e_hat, conv = dec.decode_batch(syndrome)
        dt = time.perf_counter() - t0
        e_diff = (e_hat.astype(np.int64) + err) % 2
        logical = ((e_diff @ code.hz_perp.T) % 2).any(axis=1)                 # simulation.py:90-91
        results[name] = dict(num_flagged=int((1 - conv.astype(np.int64)).sum()), num_logical=int(logical.sum()),
                             ler=float(logical.mean()), elapsed=dt)
        if verbose:
            print(f"{name}: num flagged error {results[name]['num_flagged']}")
            print(f"{name}: num logical error {results[name]['num_logical']}/{num_shots}, LER {results[name]['ler']}")
            print("Elapsed time:", dt)
    return results
