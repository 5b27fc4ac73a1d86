#!/usr/bin/env python3
"""Row-strip run on N GPUs vs the same domain on one GPU (launched with torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py [scheme] [precision] [rows] [cols] [iters]

Every rank steps its strip through the C ABI (NCCL halo exchange + dt all-reduce inside the
library); rank 0 also runs the whole domain on its own GPU and compares, bit for bit.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from hipims_ocl_b200 import executor as hx, strips
from tests.helpers import add_standard_boundaries, dtype_of, make_cfg, scenario


def main():
    scheme = sys.argv[1] if len(sys.argv) > 1 else "godunov"
    precision = sys.argv[2] if len(sys.argv) > 2 else "double"
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    cols = int(sys.argv[4]) if len(sys.argv) > 4 else 192
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 60
    bdy = sys.argv[6] if len(sys.argv) > 6 else "cells"
    options = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    exchange = sys.argv[8] if len(sys.argv) > 8 else "nccl"       # "nccl" | "peer" (row strips over peer memory, no NCCL in the loop)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = dtype_of(precision)
    bed, st, man = scenario("valley", rows, cols, dt) if rows != cols else scenario("valley", rows, cols, dt)
    cfg_full = make_cfg(scheme, precision, rows, cols)
    strip = strips.make_strip(rows, world, rank, scheme)
    cfg = cfg_full.with_(rows=strip.rows)
    ex = hx.Executor(local)
    sim = hx.CudaScheme(ex, cfg, options=options, global_rows=rows, row_offset=strip.row_offset, halo_south=strip.halo_south,
                        halo_north=strip.halo_north)
    sl = strip.local_slice()
    if exchange == "peer":
        sim.upload(st[sl], bed[sl], man[sl])
        sim.sync()
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.peer_export())
        sim.attach_peers(rank, world, blobs)
    else:
        ids = [hx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.attach_comm(ids[0], rank, world)
        sim.upload(st[sl], bed[sl], man[sl])
    add_standard_boundaries(sim, cfg_full, bdy)   # global cell ids; the library keeps the ones it holds
    sim.set_target(1e6)
    sim.iterate(iters)
    mine = sim.download()[strip.owned_local_slice()]
    stats = sim.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (strip.row_offset, mine, stats))
    ok = True
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])], axis=0)
        ref = hx.CudaScheme(ex, cfg_full, options=options)
        ref.upload(st, bed, man)
        add_standard_boundaries(ref, cfg_full, bdy)
        ref.set_target(1e6)
        ref.iterate(iters)
        want = ref.download()
        same_state = np.array_equal(full, want, equal_nan=True)
        same_clock = all(g[2] == ref.stats() for g in gathered)
        err = float(np.nan_to_num(np.abs(full - want)).max())
        print("multigpu_check %s %s %dx%d x%d ranks, %d iterations, bdy=%s options=%d exchange=%s: state %s (max diff %.3e), clocks %s, t=%.6f" % (
            scheme, precision, rows, cols, world, iters, bdy, options, exchange, "IDENTICAL" if same_state else "DIFFERENT", err,
            "IDENTICAL" if same_clock else "DIFFERENT", ref.stats()["time"]))
        if not same_state:
            d = np.abs(full - want).max(axis=2)
            ys, xs = np.nonzero(d > 0)
            print("multigpu_check detail: %d cells differ, rows %s cols %s, strip height %d" % (
                len(ys), np.unique(ys)[:24].tolist(), np.unique(xs)[:24].tolist(), rows // world))
        ok = same_state and same_clock
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
