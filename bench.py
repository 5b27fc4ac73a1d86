#!/usr/bin/env python3
"""bench.py -- throughput of the explicit cell-update hot path in cell-updates/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one iteration of the scheme over every cell of the domain: boundaries -> cell update
-> CFL reduction -> time advance (one CSchemeGodunov::scheduleIteration of the reference).

Default workload (BASELINE.json configs; the north star is the fp64 MUSCL-Hancock/HLLC step):
  N = 1   pluvial16384  configs[2]: time-varying uniform rain on a 16384^2 fractal DEM, MUSCL-Hancock + MINMOD fp64,
                        Manning friction -- the largest single-GPU configuration (268 M cells, 21.5 GB resident)
  N > 1   river32768    configs[4]: 32768 columns x 4096*N rows river / tidal valley with imposed discharge and level
                        cells, MUSCL-Hancock fp64, one 4096-row strip per GPU (weak scaling; N = 8 is the 32768^2
                        domain), NCCL halo exchange + dt all-reduce
Others with --workload: dambreak4096 (configs[1], Godunov fp64; -f32, -mh, -inertial variants), radar16384
(configs[3]), newcastle (configs[0] shape).  At N = 1 the 4096^2 dam-break family is also run briefly and reported
under "variants".

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through
the C ABI with HOST (pinned) buffers: upload of the whole domain, K iterations, read-back of the
clock and of the final cell states, all inside the timed region.  Everything one-off -- CUDA graph
capture (with the NCCL exchange inside), NCCL connection set-up -- is done before the warm-up steps,
whatever --warmup is.  At N > 1 the line also carries `phases` (per-phase device times of the strip
iteration) and `strip_parity` (strips against one GPU, bit for bit, on a slab of the same workload).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU legs (cpu_baseline, --impl reference) use every host thread through OpenMP.  torchrun exports OMP_NUM_THREADS=1
# to its workers, and libgomp reads the variable when it is first loaded: set it before anything can load it.
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hipims_ocl_b200 import SchemeConfig  # noqa: E402
from hipims_ocl_b200 import scenarios as sc  # noqa: E402

WORKLOADS = {
    "dambreak4096": dict(scheme="godunov", precision="double", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "dambreak4096-f32": dict(scheme="godunov", precision="single", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "dambreak4096-mh": dict(scheme="muscl-hancock", precision="double", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "dambreak4096-mh-f32": dict(scheme="muscl-hancock", precision="single", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "dambreak4096-inertial": dict(scheme="inertial", precision="double", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "dambreak4096-inertial-f32": dict(scheme="inertial", precision="single", cols=4096, rows_per_gpu=4096, scenario="dambreak"),
    "pluvial4096-mh": dict(scheme="muscl-hancock", precision="double", cols=4096, rows_per_gpu=4096, scenario="pluvial"),
    "pluvial4096": dict(scheme="godunov", precision="double", cols=4096, rows_per_gpu=4096, scenario="pluvial"),
    # configs[0] shape (the DEM itself needs an HFA reader): 342 x 195 at 2 m, rain 70 mm/h + losses 12 mm/h
    "newcastle": dict(scheme="godunov", precision="double", cols=342, rows_per_gpu=195, scenario="pluvial", delta=2.0,
                      boundaries="rain+loss"),
    # configs[2]: uniform time-varying rain on a 16384^2 fractal DEM, MUSCL-Hancock fp64
    "pluvial16384": dict(scheme="muscl-hancock", precision="double", cols=16384, rows_per_gpu=16384, scenario="pluvial",
                         boundaries="rain-series"),
    # configs[3]: 16384^2, inertial fp32, gridded radar rain + 1024 point volume sources
    "radar16384": dict(scheme="inertial", precision="single", cols=16384, rows_per_gpu=16384, scenario="pluvial",
                       boundaries="radar+sewers"),
    # configs[4]: 32768 columns x 4096 rows per GPU, MUSCL-Hancock fp64, imposed discharge (west) and level (east)
    "river32768": dict(scheme="muscl-hancock", precision="double", cols=32768, rows_per_gpu=4096, scenario="valley",
                       boundaries="river"),
}


def attach_boundaries(sim, w, cols, total_rows):
    """Boundary sets of the BASELINE configs; cell ids are global."""
    kind = w.get("boundaries")
    if not kind:
        return
    from hipims_ocl_b200 import config as hc
    if kind == "rain+loss":
        sim.add_uniform(hc.UNIFORM_LOSS_RATE, [0.0, 1.0e8], [12.0, 12.0])
        sim.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 3600.0, 7200.0, 10800.0], [70.0, 70.0, 0.0, 0.0])
    elif kind == "rain-series":
        sim.add_uniform(hc.UNIFORM_RAIN_INTENSITY, [0.0, 1800.0, 3600.0, 5400.0], [50.0, 100.0, 0.0, 0.0])
    elif kind == "radar+sewers":
        rng = np.random.default_rng(7)
        gr, gc = -(-total_rows // 256), -(-cols // 256)
        sim.add_gridded(hc.GRIDDED_RAIN_INTENSITY, 300.0, 256.0, 0.0, 0.0, rng.uniform(0.0, 80.0, size=(13, gr, gc)))
        pts = rng.integers(1, [total_rows - 1, cols - 1], size=(1024, 2))
        ids = sorted(set(int(y) * cols + int(x) for y, x in pts))
        # evenly spaced like every series the reference's kernels can index (CLBoundaries.clc:43-52)
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_VOLUME, ids,
                     [[600.0 * i, 0.0, q, 0.0] for i, q in enumerate([0.0, 2.0, 0.0] + [0.0] * 22)])
    elif kind == "river":
        # one river per 4096-row strip (see make_inputs): every strip carries the same forcing
        period = w["rows_per_gpu"]
        band = max(4, period // 64)
        mids = [k * period + period // 2 for k in range(max(1, total_rows // period))]
        west = [y * cols + 1 for mid in mids for y in range(mid - band, mid + band)]
        ts = np.array([[600.0 * i, 0.0, 5000.0 if i else 0.0, 0.0] for i in range(150)])   # evenly spaced, 25 h
        ts[:, 2] /= len(west)
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_DISCHARGE, west, ts)
        east = [y * cols + cols - 2 for mid in mids for y in range(mid - band, mid + band)]
        t = np.arange(0.0, 44700.0 * 2, 447.0)
        tide = np.stack([t, 3.0 + 2.0 * np.sin(2 * np.pi * t / 44700.0), 0 * t, 0 * t], axis=1)
        sim.add_cell(hc.DEPTH_IS_FSL, hc.DISCHARGE_IGNORE, east, tide)
    else:
        raise ValueError(kind)


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_record(workload):
    """Figures of the dominant kernel from the committed ncu capture (profiles/kernels.json), if any."""
    path = os.path.join(ROOT, "profiles", "kernels.json")
    if os.path.exists(path):
        try:
            rec = json.load(open(path)).get(workload)
            return rec if isinstance(rec, dict) else None
        except Exception:
            return None
    return None


def measured_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    rec = ncu_record(workload)
    return rec.get("dram_bytes_per_launch") if rec else None


def algorithmic_bytes_per_cell(cfg):
    """read eta, eta_max, qx, qy, zb, n and write eta, eta_max, qx, qy."""
    reals = 10 if cfg.friction else 9
    return reals * cfg.real_bytes


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def position_noise(y, x):
    """Deterministic noise in [-1, 1) from the GLOBAL cell position (integer hash), so that a strip generated on its own
    agrees row for row with the same rows of the whole domain (halo rows included)."""
    h = (y.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (x.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
    h ^= h >> np.uint64(29)
    h *= np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(32)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def make_inputs(w, rows, cols, dtype, row_offset=0, total_rows=None):
    """Host arrays of one strip (rows [row_offset, row_offset+rows) of a total_rows-tall domain).  Every value is a
    function of the global cell position only."""
    total_rows = rows if total_rows is None else total_rows
    if w["scenario"] == "dambreak":
        y, x = np.mgrid[row_offset:row_offset + rows, 0:cols]
        r2 = (x - cols / 2.0 + 0.5) ** 2 + (y - total_rows / 2.0 + 0.5) ** 2
        radius = min(cols, total_rows) / 8.0
        bed = np.zeros((rows, cols))
        depth = np.where(r2 < radius * radius, 10.0, 1.0)
        return bed.astype(dtype), sc.make_states(bed, depth, dtype=dtype), np.full((rows, cols), 0.03, dtype=dtype)
    if w["scenario"] == "pluvial":
        tile = sc.fractal_dem(2048, 2048, 20260817)
        ys = np.arange(row_offset, row_offset + rows) % 2048
        xs_i = np.arange(cols) % 2048
        bed = tile[ys][:, xs_i]
        xs = np.arange(cols)[None, :] * 0.002
        bed = sc.round4(bed + xs)
        level = np.quantile(tile, 0.3)
        depth = sc.round4(np.maximum(level - bed, 0.0))
        return bed.astype(dtype), sc.make_states(bed, depth, dtype=dtype), np.full((rows, cols), 0.035, dtype=dtype)
    if w["scenario"] == "valley":
        gy = np.arange(row_offset, row_offset + rows, dtype=np.int64)[:, None]
        gx = np.arange(cols, dtype=np.int64)[None, :]
        x = gx.astype(np.float64)
        # the valley repeats every rows_per_gpu rows (one west->east river per strip), so that the weak-scaling runs
        # give every rank the same wet/dry mix as the single-GPU strip instead of one wet rank and seven dry ones
        period = float(w["rows_per_gpu"])
        y = np.mod(gy.astype(np.float64), period)
        mid, width = period / 2.0, max(8.0, period / 16.0)
        bed = 0.001 * (cols - x) + 20.0 * (1.0 - np.exp(-(((y - mid) / width) ** 2))) + 1.0
        bed = sc.round4(bed + 0.5 * position_noise(np.broadcast_to(gy, (rows, cols)), np.broadcast_to(gx, (rows, cols))))
        depth = sc.round4(np.maximum(0.001 * (cols - x) + 3.0 - bed, 0.0))
        return bed.astype(dtype), sc.make_states(bed, depth, dtype=dtype), np.full((rows, cols), 0.03, dtype=dtype)
    raise ValueError(w["scenario"])


def cfg_for(w, rows, cols):
    return SchemeConfig(scheme=w["scheme"], precision=w["precision"], rows=rows, cols=cols, delta=w.get("delta", 1.0),
                        end_time=1.0e7, friction=True)


# ------------------------------------------------------------------------------------------------
# CPU legs: the reference's own kernels (oracle/_ref) when they were built, else the oracle port.
# ------------------------------------------------------------------------------------------------
def cpu_backend(cfg):
    from oracle import cpu_sim
    if os.path.exists(cpu_sim.ref_library_path(cfg)):
        return "ref", "reference"
    return "oracle", "port"


def cpu_rate(w, n, steps, warmup, threads=None):
    """cell-updates/s of the CPU implementation on an n x n crop of the workload, on all host threads (set explicitly:
    torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    threads = threads or (os.cpu_count() or 1)
    from oracle import cpu_sim
    cfg = cfg_for(w, n, n)
    backend, kind = cpu_backend(cfg)
    dtype = np.float64 if cfg.precision == "double" else np.float32
    bed, st, man = make_inputs(w, n, n, dtype)
    sim = cpu_sim.CpuSim(backend, cfg, threads=threads)
    sim.upload(st, bed, man)
    attach_boundaries(sim, w, n, n)
    sim.set_target(1.0e7)
    sim.iterate(warmup)
    t0 = time.perf_counter()
    sim.iterate(steps)
    dt = time.perf_counter() - t0
    sim.close()
    return n * n * steps / dt, dt, kind


def cpu_baseline(w, budget_s=12.0):
    cores = os.cpu_count() or 1
    n = min(w["cols"], 2048)
    rate, _, kind = cpu_rate(w, n, 3, 1)          # calibrate at the sample size, then fill the budget
    steps = int(max(3, min(2000, budget_s * rate / (n * n))))
    rate, dt, kind = cpu_rate(w, n, steps, 1)
    return {"value": rate, "unit": "cell-updates/s", "cores": cores, "kind": kind,
            "sample": "%d steps of a %dx%d crop of the workload, all %d host threads (OpenMP), %.1f s" % (steps, n, n, cores, dt)}


def config_dict(name, w, world, options=0):
    """The `config` object of the JSON line -- built by ONE function for both arms, so that they are equal."""
    cols, rows_own = w["cols"], w["rows_per_gpu"]
    rb = 8 if w["precision"] == "double" else 4
    halo = 2 if w["scheme"] == "muscl-hancock" else 1
    rows_held = rows_own + (2 * halo if world > 2 else halo if world == 2 else 0)
    return {"workload": name, "scheme": w["scheme"], "riemann_solver": "hllc", "cols": cols, "rows": rows_own * world,
            "rows_per_gpu": rows_own, "friction": True, "timestep": "cfl 0.5",
            "boundaries": w.get("boundaries", "none"),
            "decomposition": "row strips x%d" % world if world > 1 else "single domain",
            "l2": "state planes %.0f MB per GPU exceed the 126 MB L2; no flush needed" % (10 * cols * rows_held * rb / 1e6),
            "options": options}


def run_reference_arm(args, w, name):
    """The reference's own kernels (oracle/_ref, compiled from the reference's .clc sources; else the oracle port) on all
    host threads; each step is one iteration over a bounded crop of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    rate, _, kind = cpu_rate(w, 512, 2, 1)
    # size the crop so that the whole run takes about two minutes at most
    n = int(min(w["cols"], max(256, (120.0 * rate / total) ** 0.5))) // 256 * 256
    n = max(256, n)
    rate, dt, kind = cpu_rate(w, n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": rate, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if w["precision"] == "double" else "f32",
        "data": "synthetic", "config": config_dict(name, w, max(1, args.gpus), args.options),
        "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": cores, "kind": kind,
                         "sample": "each step = one iteration over a %dx%d crop of %s on %d host threads (OpenMP)" % (n, n, name, cores)},
        "e2e": {"value": rate, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def issue_roofline(name, value_per_gpu, sm_mhz):
    """Second roofline: these kernels are bound by instruction issue, not DRAM.  From the committed opcode
    histogram of the same kernel: cycles per 32 cell-updates >= 2.2 * N_fp64 + N_other on each of the 592 schedulers."""
    rec = ncu_record(name)
    if not rec or "fp64_instructions_per_32_cells" not in rec:
        return None
    n64, nall = rec["fp64_instructions_per_32_cells"], rec["warp_instructions_per_32_cells"]
    ghz = (sm_mhz or 1965.0) * 1e-3
    peak = 592 * ghz * 1e9 * 32 / (2.2 * n64 + (nall - n64))
    return {"bound": "fp64-issue", "achieved": value_per_gpu, "peak": peak, "unit": "cell-updates/s", "frac": value_per_gpu / peak,
            "model": "592 schedulers x SM clock x 32 / (2.2 x fp64 + other warp instructions per 32 cell-updates)",
            "fp64_instructions_per_32_cells": n64, "warp_instructions_per_32_cells": nall,
            "source": "committed ncu capture (profiles/%s)" % rec.get("profile", "kernels.json")}


def strip_parity_check(hx, ex, dist, w, rank, world, options, exchange="nccl"):
    """Strips against ONE GPU, bit for bit: a 4096-column slab of the same workload with 256 rows per rank, 10 iterations
    through the same library calls as the timed run (NCCL halo exchange + dt all-reduce); rank 0 also runs the whole slab
    on its own GPU and compares states and clocks (what CDomainLink's exchange must guarantee,
    src/Domain/Links/CDomainLink.cpp:168-270)."""
    from hipims_ocl_b200 import strips
    cols, rows_own, iters = min(w["cols"], 4096), 256, 10
    ws = dict(w, cols=cols, rows_per_gpu=rows_own)
    total = rows_own * world
    cfg_full = cfg_for(ws, total, cols)
    dtype = np.float64 if cfg_full.precision == "double" else np.float32
    strip = strips.make_strip(total, world, rank, cfg_full.scheme)
    cfg = cfg_full.with_(rows=strip.rows)
    bed, st, man = make_inputs(ws, strip.rows, cols, dtype, row_offset=strip.row_offset - strip.halo_south, total_rows=total)
    sim = hx.CudaScheme(ex, cfg, options=options, global_rows=total, row_offset=strip.row_offset,
                        halo_south=strip.halo_south, halo_north=strip.halo_north)
    if exchange == "peer":
        sim.upload(st, bed, man)
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.peer_export())
        sim.attach_peers(rank, world, blobs)
    else:
        ids = [hx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.attach_comm(ids[0], rank, world)
        sim.upload(st, bed, man)
    attach_boundaries(sim, ws, cols, total)
    sim.set_target(1.0e7)
    sim.iterate(iters)
    mine = sim.download()[strip.owned_local_slice()]
    stats = sim.stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (strip.row_offset, mine, stats))
    result = None
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])], axis=0)
        bed, st, man = make_inputs(ws, total, cols, dtype, row_offset=0, total_rows=total)
        ref = hx.CudaScheme(ex, cfg_full, options=options)
        ref.upload(st, bed, man)
        attach_boundaries(ref, ws, cols, total)
        ref.set_target(1.0e7)
        ref.iterate(iters)
        want, want_stats = ref.download(), ref.stats()
        same = np.array_equal(full, want) and all(g[2] == want_stats for g in gathered)
        changed = not np.array_equal(want, st)
        result = {"verdict": "identical" if same and changed else "DIFFERENT", "slab": "%d x %d" % (cols, total),
                  "iterations": iters, "max_abs_diff": float(np.abs(full - want).max()), "time": want_stats["time"],
                  "wet_cells": int(((want[..., 0] - bed) > 1e-10).sum())}
        ref.close()
    dist.barrier()
    sim.close()
    return result


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: pluvial16384 (configs[2]) on one GPU, river32768 (configs[4]) on several")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-strip-parity", action="store_true")
    ap.add_argument("--options", type=int, default=0, help="HP_OPT_* bit mask")
    ap.add_argument("--exchange", choices=["nccl", "peer"], default=os.environ.get("HIPIMS_BENCH_EXCHANGE", "nccl"),
                    help="row strips at N > 1: NCCL send/recv + all-reduce (default) or the peer-memory kernel (hp_scheme_attach_peers)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload or ("pluvial16384" if max(args.gpus, world) == 1 else "river32768")
    w = WORKLOADS[name]

    if args.impl == "reference":
        run_reference_arm(args, w, name)
        return

    import torch
    from hipims_ocl_b200 import executor as hx

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA executor has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cols, rows_own = w["cols"], w["rows_per_gpu"]
    total_rows = rows_own * world
    halo = 2 if w["scheme"] == "muscl-hancock" else 1
    hs, hn = (halo if rank > 0 else 0), (halo if rank < world - 1 else 0)
    rows = rows_own + hs + hn
    cfg = cfg_for(w, rows, cols)
    dtype = np.float64 if cfg.precision == "double" else np.float32
    bed, st, man = make_inputs(w, rows, cols, dtype, row_offset=rank * rows_own - hs, total_rows=total_rows)

    ex = hx.Executor(local_rank)
    sim = hx.CudaScheme(ex, cfg, options=args.options, global_rows=total_rows, row_offset=rank * rows_own, halo_south=hs,
                        halo_north=hn)
    peer = world > 1 and args.exchange == "peer"
    if world > 1 and not peer:
        ids = [hx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.attach_comm(ids[0], rank, world)
    attach_boundaries(sim, w, cols, total_rows)

    # pinned host buffers for the end-to-end leg; the cell states are read back into the array they were uploaded from,
    # as the reference does with its host copy of the domain (CDomain::pCellStates)
    t_st = torch.from_numpy(st).pin_memory()
    del st
    t_bed = torch.from_numpy(bed).pin_memory()
    t_man = torch.from_numpy(man).pin_memory()
    del bed, man
    h2d = t_st.numel() * t_st.element_size() + t_bed.numel() * t_bed.element_size() + t_man.numel() * t_man.element_size()
    d2h = t_st.numel() * t_st.element_size() + 72

    def reset():
        sim.upload_ptrs(t_st.data_ptr(), t_bed.data_ptr(), t_man.data_ptr())
        sim.set_clock(0.0, cfg.initial_dt, 0.0)
        sim.reset_counters()
        sim.set_target(1.0e7)
        sim.sync()

    # ---- one-off work, before the warm-up steps and whatever --warmup is --------------------------
    # CUDA graphs (with the NCCL exchange captured inside for small strips) are built and uploaded here, and 36
    # untimed iterations replay each of them twice and let NCCL set up its peer connections.
    reset()
    if peer:
        # row strips over peer memory instead of NCCL (hp_scheme_attach_peers; a rendezvous, after the upload)
        blobs = [None] * world
        dist.all_gather_object(blobs, sim.peer_export())
        sim.attach_peers(rank, world, blobs)
    sim.prepare_graphs()
    sim.iterate(36, sync=True)
    # Spin-up: the timed steps must see the workload, not its dry initial state.  Rain is applied about once per
    # simulated second (hydrological accumulator), so workloads with rain run until it has fallen twice;
    # the spun-up state is what both timed legs step (it is read back into the host buffer the e2e leg uploads).
    spinup = 36
    if "rain" in str(w.get("boundaries")) or "radar" in str(w.get("boundaries")):
        while sim.raw_stats().time < 2.2 and spinup < 1200:
            sim.iterate(32, sync=True)
            spinup += 32
    sim.download_ptr(t_st.data_ptr())
    spun_up = sim.raw_stats()
    barrier()

    # ---- device-resident throughput ------------------------------------------------------------
    sim.iterate(args.warmup, sync=True)
    launches0 = sim.raw_stats().kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ex.timer_start()
    sim.iterate(args.steps, sync=False)
    ms = ex.timer_stop()
    barrier()
    clocks = sampler.stop()
    stats = sim.raw_stats()
    launches = stats.kernel_launches - launches0
    ms_t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    cells_total = cols * total_rows
    value = cells_total * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    sim.sync()
    barrier()
    t0 = time.perf_counter()
    ex.timer_start()
    sim.upload_ptrs(t_st.data_ptr(), t_bed.data_ptr(), t_man.data_ptr())
    sim.iterate(args.steps, sync=False)
    final = sim.raw_stats()                      # D2H of the clock record (synchronous)
    sim.download_ptr(t_st.data_ptr())            # D2H of the cell states (synchronous)
    e2e_ms = ex.timer_stop()
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_t = torch.tensor([max(e2e_ms, e2e_wall)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = cells_total * args.steps / (float(e2e_t.item()) * 1e-3)
    assert np.isfinite(t_st.numpy()[..., 0]).all() and final.batch_successful > 0

    # ---- multi-GPU: where the strip iteration spends its time, and strips against one GPU ------------
    phases, parity, single = None, None, None
    if world > 1 and peer:
        if not args.no_strip_parity:
            parity = strip_parity_check(hx, ex, dist, w, rank, world, args.options, "peer")
        dist.barrier()
    elif world > 1:
        sim.strip_timing(True)
        sim.iterate(4, sync=True)
        sim.strip_timing(True)                   # clears the accumulators: the first timed-mode iterations settle
        sim.iterate(12, sync=True)
        ph, n_ph = sim.strip_phases()
        sim.strip_timing(False)
        ph_t = torch.tensor([ph[k] for k in hx.STRIP_PHASES], dtype=torch.float64, device="cuda")
        ph_max = ph_t.clone()
        dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
        phases = {"unit": "ms per iteration", "iterations": n_ph, "rank0": ph,
                  "max_over_ranks": dict(zip(hx.STRIP_PHASES, [float(v) for v in ph_max.tolist()])),
                  "note": "direct launches with one host sync per iteration (diagnostic mode); small strips: interior_rows is "
                          "the whole step and allreduce is the halo exchange + all-reduce NCCL group"}
        if not args.no_strip_parity:
            parity = strip_parity_check(hx, ex, dist, w, rank, world, args.options)
        # the same per-GPU workload on ONE GPU (rank 0's, no communicator), same steps and warm-up: the denominator of the
        # weak-scaling efficiency measured in the same run (the N = 1 default of this script is another workload)
        if rank == 0:
            sim.sync()
            b1, s1, m1 = make_inputs(w, rows_own, cols, dtype, row_offset=0, total_rows=rows_own)
            one = hx.CudaScheme(ex, cfg_for(w, rows_own, cols), options=args.options)
            one.upload(s1, b1, m1)
            del b1, s1, m1
            attach_boundaries(one, w, cols, rows_own)
            one.set_target(1.0e7)
            one.prepare_graphs()
            one.iterate(36 + args.warmup, sync=True)
            ex.timer_start()
            one.iterate(args.steps, sync=False)
            ms1 = ex.timer_stop()
            one.close()
            single = {"value": cols * rows_own * args.steps / (ms1 * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms1 / args.steps,
                      "note": "one %d-row strip of the same workload on one GPU without a communicator, same steps; "
                              "value / (n_gpus x this) is the weak-scaling efficiency" % rows_own}
        dist.barrier()

    if rank == 0:
        peak, peak_src = peak_hbm()
        abytes = algorithmic_bytes_per_cell(cfg)
        kernel_s = ms * 1e-3 / args.steps          # one step per iteration on this rank (edge + interior launches on large strips)
        achieved = abytes * cols * rows_own / kernel_s / 1e9
        rec = ncu_record(name)
        traffic = rec["dram_bytes_per_cell"] * cols * rows_own if rec and "dram_bytes_per_cell" in rec else None
        line = {
            "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if cfg.precision == "double" else "f32", "data": "synthetic",
            "config": config_dict(name, w, world, args.options),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_source": ("committed ncu capture (profiles/%s): %.1f dram bytes per cell-update x the cells one "
                                            "launch covers on this GPU" % (rec.get("profile", "kernels.json"), rec["dram_bytes_per_cell"]))
                                           if traffic else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_cell": abytes, "kernel_ms": kernel_s * 1e3,
                         "kernel": rec.get("kernel") if rec else None,
                         # the kernels are bound by instruction issue / the fp64 pipe, not by DRAM:
                         # figures of the same kernel FROM THE COMMITTED CAPTURE, not measured in this run
                         "ncu_from_committed_capture": rec},
            "roofline_issue": issue_roofline(name, value / world, clocks.get("sm_mhz")),
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps,
                    "note": "upload of the whole domain + K iterations + clock and state read-back, amortised per step; with "
                            "K = %d the host<->device copies are %.0f %% of this leg" %
                            (args.steps, 100.0 * max(0.0, 1.0 - (ms_max / max(float(e2e_t.item()), 1e-9))))},
            "gpu_launches": int(launches), "clocks": clocks,
            "sim": {"time": final.time, "timestep": final.timestep, "successful": final.batch_successful,
                    "spinup_iterations": spinup, "spinup_time": spun_up.time,
                    "note": "both timed legs step the spun-up state (after the rain has fallen, where the workload has rain)"},
        }
        if phases:
            line["phases"] = phases
        if parity:
            line["strip_parity"] = parity["verdict"]
            line["strip_parity_detail"] = parity
        if single:
            line["single_gpu_same_workload"] = single
        if world > 1:
            line["exchange"] = ("peer memory: one kernel stores the edge rows into the neighbours' halo rows over NVLink, exchanges the "
                                "wave-speed maximum through mailboxes and runs the time controller (hp_scheme_attach_peers)") if peer else \
                               "NCCL: ncclSend/ncclRecv of the halo rows + ncclAllReduce(max) of the wave speed (hp_scheme_attach_comm)"
        if not args.no_variants and world == 1 and args.workload is None:
            sim.close()
            del t_st, t_bed, t_man
            line["variants"] = run_variants(hx, ex, args)
        if not args.no_cpu_baseline and world == 1:
            sim.close()
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        sim.close()                  # the NCCL communicator of the strip goes before the process group
        dist.destroy_process_group()
    if parity and parity["verdict"] != "identical":
        raise SystemExit("strip parity FAILED: %r" % (parity,))


def run_variants(hx, ex, args):
    """Short device-resident runs of configs[1] (4096^2 circular dam break, Godunov fp64 / fp32) and of the other schemes
    on the same domain."""
    out = {}
    # "+march": the same workload on the Godunov marching kernels (HP_OPT_MARCH_GODUNOV: one column per lane in fp64, two
    # in fp32) instead of the scheme's default tile kernel
    for key in ("dambreak4096", "dambreak4096-f32", "dambreak4096-mh", "dambreak4096-mh-f32", "dambreak4096-inertial",
                "dambreak4096-inertial-f32", "dambreak4096+march", "dambreak4096-f32+march"):
        name, _, extra = key.partition("+")
        # a side measurement must never take the headline down with it: a failure is recorded in its place
        try:
            w = WORKLOADS[name]
            cfg = cfg_for(w, w["rows_per_gpu"], w["cols"])
            dtype = np.float64 if cfg.precision == "double" else np.float32
            bed, st, man = make_inputs(w, cfg.rows, cfg.cols, dtype)
            sim = hx.CudaScheme(ex, cfg, options=args.options | (hx.OPT_MARCH_GODUNOV if extra == "march" else 0))
            try:
                sim.upload(st, bed, man)
                sim.set_target(1.0e7)
                steps = max(20, args.steps // 2)
                sim.prepare_graphs()
                sim.iterate(max(args.warmup, 18), sync=True)
                ex.timer_start()
                sim.iterate(steps, sync=False)
                ms = ex.timer_stop()
            finally:
                sim.close()
            peak, _ = peak_hbm()
            rate = cfg.cells * steps / (ms * 1e-3)
            out[key] = {"value": rate, "unit": "cell-updates/s", "steps": steps,
                        "roofline_frac": rate * algorithmic_bytes_per_cell(cfg) / 1e9 / peak}
        except Exception as exc:                                            # noqa: BLE001 -- reported, not swallowed
            out[key] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    return out


if __name__ == "__main__":
    main()
