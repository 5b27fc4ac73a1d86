#!/usr/bin/env python3
"""bench.py -- throughput of the explicit cell-update hot path in cell-updates/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one iteration of the scheme over every cell of the domain: boundaries -> cell update
-> CFL reduction -> time advance (one CSchemeGodunov::scheduleIteration of the reference).

Workloads (BASELINE.json configs):
  dambreak4096   configs[1]: circular dam break, 4096 x 4096 flat DEM, Godunov/HLLC, CFL timestep,
                 friction on (default at N=1; at N>1 the domain is 4096 x 4096*N, one 4096-row strip
                 per GPU -- weak scaling)
  pluvial16384   configs[2]: uniform rain on a 16384^2 fractal DEM, MUSCL-Hancock fp64 (--workload)
  river32768     configs[4]: 32768 columns x 4096*N rows river valley, MUSCL-Hancock fp64 (--workload)

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through
the C ABI with HOST (pinned) buffers: upload of the whole domain, K iterations, read-back of the
clock and of the final cell states, all inside the timed region.
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
        sim.add_cell(hc.DEPTH_IGNORE, hc.DISCHARGE_IS_VOLUME, ids, [[0.0, 0.0, 0.0, 0.0], [600.0, 0.0, 2.0, 0.0], [1200.0, 0.0, 0.0, 0.0], [1.0e6, 0.0, 0.0, 0.0]])
    elif kind == "river":
        # one river per 4096-row strip (see make_inputs): every strip carries the same forcing
        period = w["rows_per_gpu"]
        band = max(4, period // 64)
        mids = [k * period + period // 2 for k in range(max(1, total_rows // period))]
        west = [y * cols + 1 for mid in mids for y in range(mid - band, mid + band)]
        ts = np.array([[0.0, 0.0, 0.0, 0.0], [600.0, 0.0, 5000.0, 0.0], [1.0e6, 0.0, 5000.0, 0.0]])
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


def make_inputs(w, rows, cols, dtype, row_offset=0, total_rows=None):
    """Host arrays of one strip (rows [row_offset, row_offset+rows) of a total_rows-tall domain)."""
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
        reps = (-(-rows // 2048), -(-cols // 2048))
        bed = np.tile(tile, reps)[:rows, :cols]
        xs = np.arange(cols)[None, :] * 0.002
        bed = sc.round4(bed + xs)
        level = np.quantile(tile, 0.3)
        depth = sc.round4(np.maximum(level - bed, 0.0))
        return bed.astype(dtype), sc.make_states(bed, depth, dtype=dtype), np.full((rows, cols), 0.035, dtype=dtype)
    if w["scenario"] == "valley":
        y = (np.arange(row_offset, row_offset + rows, dtype=np.float64))[:, None]
        x = np.arange(cols, dtype=np.float64)[None, :]
        # the valley repeats every rows_per_gpu rows (one west->east river per strip), so that the weak-scaling runs
        # give every rank the same wet/dry mix as the single-GPU strip instead of one wet rank and seven dry ones
        period = float(w["rows_per_gpu"])
        y = np.mod(y, period)
        mid, width = period / 2.0, max(8.0, period / 16.0)
        rng = np.random.default_rng(20260819 + row_offset)
        bed = 0.001 * (cols - x) + 20.0 * (1.0 - np.exp(-(((y - mid) / width) ** 2))) + 1.0
        bed = sc.round4(bed + 0.5 * rng.uniform(-1.0, 1.0, size=(rows, cols)))
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


def run_reference_arm(args, w, name):
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
    cfg = cfg_for(w, n, n)
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": rate, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if cfg.precision == "double" else "f32",
        "data": "synthetic", "config": {"workload": name, "scheme": cfg.scheme, "crop": "%dx%d" % (n, n)},
        "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": cores, "kind": kind,
                         "sample": "each step = one iteration over a %dx%d crop of %s on %d host threads" % (n, n, name, cores)},
        "e2e": {"value": rate, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dambreak4096", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--options", type=int, default=0, help="HP_OPT_* bit mask")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, w, args.workload)
        return

    import torch
    from hipims_ocl_b200 import executor as hx

    world = int(os.environ.get("WORLD_SIZE", "1"))
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
    if world > 1:
        ids = [hx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.attach_comm(ids[0], rank, world)
    attach_boundaries(sim, w, cols, total_rows)

    # pinned host buffers for the end-to-end leg
    t_st = torch.from_numpy(st).pin_memory()
    t_bed = torch.from_numpy(bed).pin_memory()
    t_man = torch.from_numpy(man).pin_memory()
    t_out = torch.empty_like(t_st).pin_memory()
    h2d = t_st.numel() * t_st.element_size() + t_bed.numel() * t_bed.element_size() + t_man.numel() * t_man.element_size()
    d2h = t_out.numel() * t_out.element_size() + 72

    def reset():
        sim.upload_ptrs(t_st.data_ptr(), t_bed.data_ptr(), t_man.data_ptr())
        sim.set_clock(0.0, cfg.initial_dt, 0.0)
        sim.reset_counters()
        sim.set_target(1.0e7)
        sim.sync()

    # ---- device-resident throughput ------------------------------------------------------------
    reset()
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
    sim.download_ptr(t_out.data_ptr())           # D2H of the cell states (synchronous)
    e2e_ms = ex.timer_stop()
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_t = torch.tensor([max(e2e_ms, e2e_wall)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = cells_total * args.steps / (float(e2e_t.item()) * 1e-3)
    assert np.isfinite(t_out.numpy()[..., 0]).all() and final.batch_successful > 0

    if rank == 0:
        peak, peak_src = peak_hbm()
        abytes = algorithmic_bytes_per_cell(cfg)
        kernel_s = ms * 1e-3 / args.steps          # one step kernel per iteration on this rank
        achieved = abytes * cols * rows_own / kernel_s / 1e9
        line = {
            "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if cfg.precision == "double" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scheme": cfg.scheme, "riemann_solver": "hllc", "cols": cols,
                       "rows": total_rows, "rows_per_gpu": rows_own, "friction": True, "timestep": "cfl 0.5",
                       "decomposition": "row strips x%d" % world if world > 1 else "single domain",
                       "l2": "state planes %.0f MB per rank exceed the 126 MB L2; no flush needed" %
                             (10 * cols * rows * cfg.real_bytes / 1e6),
                       "options": args.options},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": abytes, "kernel_ms": kernel_s * 1e3,
                         # the kernels are bound by instruction issue / the fp64 pipe, not by DRAM:
                         # the committed ncu capture of the same kernel says how far
                         "ncu": ncu_record(args.workload)},
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps,
                    "note": "upload of the whole domain + K iterations + clock and state read-back, amortised per step"},
            "gpu_launches": int(launches), "clocks": clocks,
            "sim": {"time": final.time, "timestep": final.timestep, "successful": final.batch_successful},
        }
        if not args.no_variants and world == 1 and args.workload == "dambreak4096":
            line["variants"] = run_variants(hx, ex, args)
        if not args.no_cpu_baseline and world == 1:
            sim.close()
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        sim.close()                  # the NCCL communicator of the strip goes before the process group
        dist.destroy_process_group()


def run_variants(hx, ex, args):
    """Short device-resident runs of the other precision / schemes on the same 4096^2 dam break."""
    out = {}
    for name in ("dambreak4096-f32", "dambreak4096-mh", "dambreak4096-mh-f32", "dambreak4096-inertial", "dambreak4096-inertial-f32"):
        w = WORKLOADS[name]
        cfg = cfg_for(w, w["rows_per_gpu"], w["cols"])
        dtype = np.float64 if cfg.precision == "double" else np.float32
        bed, st, man = make_inputs(w, cfg.rows, cfg.cols, dtype)
        sim = hx.CudaScheme(ex, cfg, options=args.options)
        sim.upload(st, bed, man)
        sim.set_target(1.0e7)
        steps = max(20, args.steps // 2)
        sim.iterate(args.warmup, sync=True)
        ex.timer_start()
        sim.iterate(steps, sync=False)
        ms = ex.timer_stop()
        peak, _ = peak_hbm()
        rate = cfg.cells * steps / (ms * 1e-3)
        out[name] = {"value": rate, "unit": "cell-updates/s", "steps": steps,
                     "roofline_frac": rate * algorithmic_bytes_per_cell(cfg) / 1e9 / peak}
        sim.close()
    return out


if __name__ == "__main__":
    main()
