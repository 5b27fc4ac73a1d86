"""Host-side logic of the row-strip decomposition on CPU: two or three gloo ranks step their strips with the
oracle kernels, exchange halo rows and all-reduce the wave-speed maximum exactly as the CUDA
executor does with NCCL (hp_executor.cu: enqueue_iteration), and must reproduce the single-domain
run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hipims_ocl_b200 import config as hc
from hipims_ocl_b200 import strips
from oracle import cpu_sim
from tests.helpers import dtype_of, make_cfg, scenario


def test_strip_geometry():
    for rows, world in ((64, 2), (67, 3), (4096 * 8, 8)):
        for scheme in ("godunov", "muscl-hancock"):
            cover = []
            for r in range(world):
                s = strips.make_strip(rows, world, r, scheme)
                cover += list(range(s.row_offset, s.row_offset + s.own_rows))
                assert s.halo_south == (0 if r == 0 else strips.halo_rows(scheme))
                assert s.halo_north == (0 if r == world - 1 else strips.halo_rows(scheme))
                assert s.first_local_row >= 0 and s.first_local_row + s.rows <= rows
            assert cover == list(range(rows))
    with pytest.raises(ValueError):
        strips.make_strip(8, 4, 0, "muscl-hancock")
    s = strips.make_strip(64, 2, 1, "godunov")
    assert strips.split_boundary_cells([5, 31 * 10 + 2, 63 * 10 + 9], 10, s) == [31 * 10 + 2, 63 * 10 + 9]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, scheme, precision, rows, cols, iters, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = dtype_of(precision)
    bed, st, man = scenario("valley", rows, cols, dt)
    cfg_full = make_cfg(scheme, precision, rows, cols)
    strip = strips.make_strip(rows, world, rank, scheme)
    cfg = cfg_full.with_(rows=strip.rows)
    k = cpu_sim.CpuSim("oracle", cfg, threads=1)
    sl = strip.local_slice()
    a, b = st[sl].copy(), st[sl].copy()
    zb, mn = bed[sl].copy(), man[sl].copy()
    faces = [np.zeros_like(a) for _ in range(4)]
    clock = np.array([0.0, cfg.initial_dt, 0.0, 1.0e6, 0.0], dtype=dt)
    counters = np.zeros(2, dtype=np.uint32)
    halo, own = strips.halo_rows(scheme), strip.owned_local_slice()
    alt = False
    for _ in range(iters):
        step_dt = clock[1]
        if scheme == hc.SCHEME_MUSCL_HANCOCK:
            k.k_mch_1st(step_dt, zb, a, faces)
            k.k_mch_2nd(step_dt, a, zb, mn, faces)
            dst = a
        else:
            src, dst = (b, a) if alt else (a, b)
            k.k_step(step_dt, zb, src, dst, mn)
        # halo exchange of the freshly written buffer: owned edge rows -> neighbour's halo rows
        ops, recv = [], {}
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(dst[own.start:own.start + halo].copy()), rank - 1))
            recv["south"] = torch.empty_like(torch.from_numpy(dst[:halo].copy()))
            ops.append(dist.P2POp(dist.irecv, recv["south"], rank - 1))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(dst[own.stop - halo:own.stop].copy()), rank + 1))
            recv["north"] = torch.empty_like(torch.from_numpy(dst[:halo].copy()))
            ops.append(dist.P2POp(dist.irecv, recv["north"], rank + 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if "south" in recv:
            dst[own.start - halo:own.start] = recv["south"].numpy()
        if "north" in recv:
            dst[own.stop:own.stop + halo] = recv["north"].numpy()
        # wave-speed maximum over the OWNED rows of the buffer the reference's reduction reads (Q1)
        red = a if (scheme == hc.SCHEME_MUSCL_HANCOCK or (cfg.quirks & hc.QUIRK_REDUCE_BUFFER_A)) else dst
        vmax = torch.tensor([k.k_reduce(np.ascontiguousarray(red[own]), np.ascontiguousarray(zb[own]), rows=strip.own_rows)],
                            dtype=torch.float64)
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        k.k_advance(clock, counters, float(vmax.item()))
        alt = not alt
    cur = a if (scheme == hc.SCHEME_MUSCL_HANCOCK or not alt) else b
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rows=cur[own], offset=strip.row_offset, clock=clock, counters=counters)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scheme,precision,world", [("godunov", "double", 2), ("muscl-hancock", "double", 2), ("inertial", "single", 2),
                                                    ("muscl-hancock", "double", 3)])      # 3 ranks: a strip with two neighbours
def test_strips_match_single_domain(tmp_path, scheme, precision, world):
    rows, cols, iters = 48, 40, 40
    mp.spawn(_worker, args=(world, _free_port(), scheme, precision, rows, cols, iters, str(tmp_path)), nprocs=world, join=True)
    parts = sorted((np.load(os.path.join(tmp_path, "rank%d.npz" % r)) for r in range(world)), key=lambda z: int(z["offset"]))
    got = np.concatenate([p["rows"] for p in parts], axis=0)
    cfg = make_cfg(scheme, precision, rows, cols)
    bed, st, man = scenario("valley", rows, cols, dtype_of(precision))
    ref = cpu_sim.CpuSim("oracle", cfg)
    ref.upload(st, bed, man)
    ref.set_target(1.0e6)
    ref.iterate(iters)
    np.testing.assert_array_equal(got, ref.download())
    s = ref.stats()
    for p in parts:
        assert float(p["clock"][0]) == s["time"] and float(p["clock"][1]) == s["timestep"]
        assert int(p["counters"][0]) == s["batch_successful"]
