#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun): the three multi-GPU modes -- left-operand sharding + NCCL all-to-all +
owner merge ("exchange", the contract design), the same with the exchange fused into the producer as P2P stores into
the owners' NVLink peer windows ("push"), and owner-computes ("owner", zero exchange) -- gathered on rank 0 and
compared with the CPU oracle.  Used by tests/test_dist_gpu.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def reference_dict(orc, f, g, limit, idx):
    r = orc.poly_mul(f, g, limit=limit, idx=idx, nthreads=os.cpu_count() or 1)
    cfs = r["cfs"] if isinstance(r["cfs"], list) else r["cfs"].tolist()
    if r["keys"].shape[1] == 1:
        return dict(zip(r["keys"][:, 0].tolist(), cfs))
    return {tuple(k): c for k, c in zip(r["keys"].tolist(), cfs)}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import orc
    from obake_b200 import _capi, dist as obd, engine, workloads as wl

    eng = engine.Engine(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    cases = []
    f, g = wl.fateman(10, 4, "u64")
    cases.append(("fateman10", f, g, None, None))
    f16, g16 = wl.fateman(16, 4, "u64")
    cases.append(("fateman16 (row kernel)", f16, g16, None, None))
    f, g = wl.sparse_mp(7, "i64")
    cases.append(("sparse7", f, g, None, None))
    cases.append(("sparse7 trunc 30", f, g, 30, None))
    cases.append(("sparse7 ptrunc", f, g, 9, [0, 3]))
    # fewer left-operand terms than ranks: some ranks have an empty slice but must still agree on the limb count
    f3 = wl.from_dict(f16.symbols, {(1, 0, 0, 0): 3, (0, 2, 0, 1): -(1 << 40)}, "u64")
    cases.append(("2 terms x fateman16 (empty slices)", f3, g16, None, None))
    fa, ga = wl.audi("u64", 4, nvars=6, n=6)
    cases.append(("audi66 d_packed f64", fa, ga, 6, None))
    ok_all = True
    px = obd.PeerExchange(eng, window_bytes=64 << 20)
    for name, f, g, limit, idx in cases:
        lin = None if f.is_f64 else max(engine.needed_limbs(f.cfs), engine.needed_limbs(g.cfs))
        hx, hy = engine.HostOperand(f, lin), engine.HostOperand(g, lin)
        ref = reference_dict(orc, f, g, limit, idx) if rank == 0 else None
        for mode in ("exchange", "push", "owner"):
            if mode == "exchange":
                p, info = obd.sharded_mul(eng, hx.view, hy.view, limit=limit, idx=idx, result_mem=_capi.OBK_MEM_HOST)
            elif mode == "push":
                p, info = px.pushed_mul(hx.view, hy.view, limit=limit, idx=idx, result_mem=_capi.OBK_MEM_HOST)
            else:
                p, info = obd.owner_mul(eng, hx.view, hy.view, limit=limit, idx=idx)
            mine = p.as_dict()
            if mode in ("exchange", "push") and p.nterms:
                a, b = obd.owner_ranges(info["nsegs"], world)[rank]
                seg = (p.keys % np.uint64(info["nsegs"])).sum(axis=1) % info["nsegs"]
                if f.tname == "u64":
                    assert ((seg >= a) & (seg < b)).all(), "rank holds a segment it does not own"
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            if rank == 0:
                full = {}
                for d in gathered:
                    assert not (set(full) & set(d)), "two ranks hold the same key"
                    full.update(d)
                if f.is_f64:
                    # relative 1e-12 per coefficient (BASELINE.json); a term that cancels exactly in one summation
                    # order may survive as a residue in the other: such a key may be missing on one side only if its
                    # value is below 1e-13 of the largest possible sum
                    scale = float(np.abs(f.cfs).max() * np.abs(g.cfs).max()) * min(f.nterms, g.nterms)
                    ok = True
                    for k in set(full) | set(ref):
                        a, b = full.get(k, 0.0), ref.get(k, 0.0)
                        if k in full and k in ref:
                            ok &= abs(a - b) <= 1e-12 * max(abs(a), abs(b)) + 1e-15 * scale
                        else:
                            ok &= abs(a - b) <= 1e-13 * scale
                else:
                    ok = full == ref
                print(f"[dist_check] {name} [{mode}]: world={world} terms={len(full)} kernel={info['partial_stats']['kernel']} "
                      f"sent_rows(rank0)={info['sent_rows']} ok={ok}", flush=True)
                ok_all &= ok
            p.free()
    px.close()
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
