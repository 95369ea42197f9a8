"""world_size-2 gloo test of the multi-GPU host logic: shard shots, decode per rank, reduce counters."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden
    from oracle.oracle import Oracle
    from slidingwindowdecoder_b200.distributed import shard_range, reduce_counters, max_over_ranks
    g = load_golden("c1_gdg_default_mt0")
    orc = Oracle(g["mat"], g["priors"])
    lo, hi = shard_range(400, rank, world)
    dec, conv, _, _ = orc.bpgdg_batch(g["synd"][lo:hi], **g["kwargs"])
    wrong = int((dec.astype(np.uint8) != g["dec"][lo:hi]).any(axis=1).sum())
    tot = reduce_counters([hi - lo, int((conv == 0).sum()), wrong])
    tmax = max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((tot, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_reduce_two_ranks():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden("c1_gdg_default_mt0")
    assert tot[0] == 400
    assert tot[1] == int((g["conv"][:400] == 0).sum())
    assert tot[2] == 0
    assert tmax == 2.0


def test_shard_range_partitions():
    from slidingwindowdecoder_b200.distributed import shard_range
    for total in (0, 1, 7, 10 ** 7):
        for world in (1, 2, 3, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
