#!/usr/bin/env python3
"""Per-rank statistics of the owner-computes product (development tool; run under torchrun)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from obake_b200 import dist as obd, engine, workloads as wl

    eng = engine.Engine(local)
    name = sys.argv[1] if len(sys.argv) > 1 else "F30"
    f, g, limit, _ = wl.named(name)
    lin = None if f.is_f64 else max(engine.needed_limbs(f.cfs), engine.needed_limbs(g.cfs))
    hx, hy = engine.HostOperand(f, lin), engine.HostOperand(g, lin)
    for it in range(4):
        p, info = obd.owner_mul(eng, hx.view, hy.view, limit=limit)
        st = p.stats
        if it == 3:
            print(f"rank {rank}/{world}: kernel {st['kernel']} ms_mul {st['ms_mul']:.3f} ms_total {st['ms_total']:.3f} prep {st['ms_prep']:.3f} "
                  f"est {st['ms_estimate']:.3f} units {st['reserved']} tiles/unit {st['table_slots']} stages {st['group_lanes']} "
                  f"terms {p.nterms} launches {st['total_launches']}", flush=True)
        p.free()
    dist.barrier()


if __name__ == "__main__":
    main()
