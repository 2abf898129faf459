"""Host-side logic of the multi-GPU path on CPU: world_size 2 and 3 over gloo.  The CUDA kernels cannot run
here, so the oracle stands in for the per-rank partial product and the owner merge (it is the checker,
allowed in tests); what is under test is the product code's partition + all-to-all plumbing
(obake_b200.dist.exchange_partials / owner_ranges) and that shard -> exchange -> merge reproduces f*g."""
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
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orc
    from obake_b200 import dist as obd, workloads as wl

    f, g = wl.sparse_mp(5, "u64")
    nsegs = 13  # the engine's prime modulus
    nx = f.nterms
    lo, hi = nx * rank // world, nx * (rank + 1) // world
    part = orc.poly_mul(f.take(slice(lo, hi)), g, nthreads=1, lacc=2)
    keys = part["keys"]
    limbs = part["cf_limbs"]
    seg = (keys[:, 0] % np.uint64(nsegs)).astype(np.int64)
    order = np.argsort(seg, kind="stable")
    keys, limbs, seg = keys[order], limbs[order], seg[order]
    ranges = obd.owner_ranges(nsegs, world)
    counts = [int(((seg >= a) & (seg < b)).sum()) for a, b in ranges]
    tk = torch.from_numpy(keys.view(np.int64).copy())
    tc = torch.from_numpy(limbs.view(np.int64).copy())
    rk, rc, rcounts = obd.exchange_partials(tk, tc, counts)
    assert rk.shape[0] == sum(rcounts) == rc.shape[0]
    # owner-side merge (checker arithmetic): exact sums by key
    rkeys = rk.numpy().view(np.uint64)[:, 0].tolist()
    vals = orc.limbs_to_int(rc.numpy().view(np.uint64))
    a, b = ranges[rank]
    acc = {}
    for k, v in zip(rkeys, vals):
        assert a <= k % nsegs < b, "received a term this rank does not own"
        acc[k] = acc.get(k, 0) + v
    acc = {k: v for k, v in acc.items() if v != 0}
    gathered = [None] * world
    dist.all_gather_object(gathered, acc)
    if rank == 0:
        full = {}
        for d in gathered:
            assert not (set(full) & set(d)), "two owners hold the same key"
            full.update(d)
        ref = orc.poly_mul(f, g, nthreads=2)
        q.put(full == dict(zip(ref["keys"][:, 0].tolist(), ref["cfs"])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_exchange_merge_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_owner_ranges_cover_all_segments():
    from obake_b200 import dist as obd

    for nsegs in (1, 2, 13, 293, 1181):
        for world in (1, 2, 3, 4, 8):
            r = obd.owner_ranges(nsegs, world)
            assert r[0][0] == 0 and r[-1][1] == nsegs
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
