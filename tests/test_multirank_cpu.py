"""world_size-2 test (gloo, CPU) of the multi-GPU plan: contiguous frame shards, noise indexed
by GLOBAL frame id, one all-reduce of the {errors, bits} counters.  The oracle stands in for
the per-rank CUDA path (the kernels themselves are covered by the -m gpu tests); what is under
test is that sharded + reduced == unsharded."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    from conftest import path_kwargs
    from oracle import oracle as O
    pkg = g.load_package()
    o = O.OraclePath(**path_kwargs("qpsk", sps=8, shaped=True))
    F = 10
    bits = np.random.default_rng(123).integers(0, 2, (F, 2 * 512), dtype=np.uint8)  # same on every rank
    sigma = o.sigma_for_ebn0(2.0)
    f0, n = pkg.shard_range(F, rank, world)
    _, dec, cnt = o.loopback(bits[f0:f0 + n], sigma=sigma, seed=0xA5A5, frame0=f0)
    t = torch.tensor(cnt, dtype=torch.int64)
    dist.all_reduce(t)  # the single collective of the path (NCCL on the GPUs)
    gathered = [None] * world
    dist.all_gather_object(gathered, (f0, n, dec))
    if rank == 0:
        q.put((t.tolist(), gathered))
    dist.destroy_process_group()


def test_sharded_loopback_equals_unsharded(orc):
    from conftest import path_kwargs
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = orc.OraclePath(**path_kwargs("qpsk", sps=8, shaped=True))
    bits = np.random.default_rng(123).integers(0, 2, (10, 2 * 512), dtype=np.uint8)
    _, dec_ref, cnt_ref = o.loopback(bits, sigma=o.sigma_for_ebn0(2.0), seed=0xA5A5, frame0=0)
    assert tuple(total) == cnt_ref and cnt_ref[0] > 0
    covered = np.zeros(10, int)
    for f0, n, dec in gathered:
        covered[f0:f0 + n] += 1
        assert np.array_equal(dec, dec_ref[f0:f0 + n])
    assert (covered == 1).all()


@pytest.mark.parametrize("total,world", [(4096, 1), (4096, 8), (10, 4), (3, 8), (0, 2), (1024, 3)])
def test_shard_range_partitions(pkg, total, world):
    spans = [pkg.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and sum(n for _, n in spans) == total
    for (a, n), (b, _) in zip(spans, spans[1:]):
        assert a + n == b
    assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_shard_channels(pkg):
    # config 5: 1024 carriers over 8 GPUs -> 128 whole channels each
    for r in range(8):
        c0, nc, f0, nf = pkg.shard_channels(1024, 16, r, 8)
        assert (c0, nc, f0, nf) == (128 * r, 128, 128 * r * 16, 128 * 16)
    with pytest.raises(ValueError):
        pkg.shard_range(10, 2, 2)
