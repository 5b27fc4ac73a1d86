"""ctypes binding of the C-ABI CUDA executor (include/hipims_cuda.h).

This is the Python face of the drop-in boundary: the same calls the reference's scheme classes
make on the OpenCL wrappers, over `libhipims_cuda.so`.  There is no fallback of any kind: if the
library has not been built (`python -m hipims_ocl_b200.build`) or no CUDA device is present,
construction raises.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import config as hc

HERE = os.path.dirname(os.path.abspath(__file__))
# HIPIMS_CUDA_LIB: development aid (tools/build_variant.py) -- another build of the same CUDA library, never a fallback
LIB_PATH = os.environ.get("HIPIMS_CUDA_LIB") or os.path.join(HERE, "libhipims_cuda.so")

OPT_STRICT_FP = 1
OPT_NO_GRAPH = 2
OPT_NO_TMA = 4
OPT_TILE_KERNELS = 8
OPT_MARCH_GODUNOV = 16
OPT_SPLIT_STRIPS = 32
OPT_NARROW_MARCH = 64
OPT_WIDE_MARCH = 128

_SCHEME_ID = {hc.SCHEME_GODUNOV: 0, hc.SCHEME_MUSCL_HANCOCK: 1, hc.SCHEME_INERTIAL: 2}


class HipimsCudaError(RuntimeError):
    pass


class HpSchemeConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("scheme", C.c_uint32), ("real_bytes", C.c_uint32),
                ("quirks", C.c_uint32), ("options", C.c_uint32), ("dynamic_timestep", C.c_uint32),
                ("friction", C.c_uint32), ("reserved0", C.c_uint32), ("cols", C.c_uint64), ("rows", C.c_uint64),
                ("delta", C.c_double), ("courant", C.c_double), ("dry_threshold", C.c_double),
                ("end_time", C.c_double), ("fixed_timestep", C.c_double), ("initial_timestep", C.c_double),
                ("global_rows", C.c_uint64), ("row_offset", C.c_uint64), ("halo_south", C.c_uint32),
                ("halo_north", C.c_uint32)]


class HpSchemeStats(C.Structure):
    _fields_ = [("time", C.c_double), ("timestep", C.c_double), ("time_hydrological", C.c_double),
                ("time_target", C.c_double), ("batch_timesteps", C.c_double), ("batch_successful", C.c_uint32),
                ("batch_skipped", C.c_uint32), ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("use_alternate", C.c_uint32), ("reserved0", C.c_uint32)]


class HpBdyUniform(C.Structure):
    _fields_ = [("entries", C.c_uint32), ("definition", C.c_uint32), ("interval", C.c_double), ("length", C.c_double)]


class HpBdyGridded(C.Structure):
    _fields_ = [("interval", C.c_double), ("resolution", C.c_double), ("offset_x", C.c_double),
                ("offset_y", C.c_double), ("entries", C.c_uint64), ("definition", C.c_uint64), ("rows", C.c_uint64),
                ("cols", C.c_uint64)]


class HpBdyCell(C.Structure):
    _fields_ = [("entries", C.c_uint64), ("interval", C.c_double), ("length", C.c_double), ("relations", C.c_uint64),
                ("def_depth", C.c_uint32), ("def_discharge", C.c_uint32)]


# every symbol include/hipims_cuda.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
ABI = {
    "hp_abi_version": (C.c_int, []),
    "hp_last_error": (C.c_char_p, []),
    "hp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hp_executor_create": (C.c_int, [C.c_int, _VP, C.POINTER(_VP)]),
    "hp_executor_destroy": (None, [_VP]),
    "hp_executor_describe": (C.c_int, [_VP, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "hp_executor_finish": (C.c_int, [_VP]),
    "hp_executor_timer_start": (C.c_int, [_VP]),
    "hp_executor_timer_stop": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "hp_scheme_create": (C.c_int, [_VP, C.POINTER(HpSchemeConfig), C.POINTER(_VP)]),
    "hp_scheme_destroy": (None, [_VP]),
    "hp_boundary_add_uniform": (C.c_int, [_VP, C.POINTER(HpBdyUniform), _VP]),
    "hp_boundary_add_gridded": (C.c_int, [_VP, C.POINTER(HpBdyGridded), _VP]),
    "hp_boundary_add_cell": (C.c_int, [_VP, C.POINTER(HpBdyCell), _VP, _VP]),
    "hp_scheme_upload_cells": (C.c_int, [_VP, _VP, _VP, _VP]),
    "hp_scheme_download_cells": (C.c_int, [_VP, _VP]),
    "hp_scheme_download_both": (C.c_int, [_VP, _VP, _VP]),
    "hp_scheme_read_rows": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP]),
    "hp_scheme_derive_raster": (C.c_int, [_VP, C.c_uint32, C.c_double, _VP]),
    "hp_scheme_write_rows": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP]),
    "hp_scheme_set_target_time": (C.c_int, [_VP, C.c_double]),
    "hp_scheme_force_timestep": (C.c_int, [_VP, C.c_double]),
    "hp_scheme_set_clock": (C.c_int, [_VP, C.c_double, C.c_double, C.c_double]),
    "hp_scheme_update_timestep": (C.c_int, [_VP]),
    "hp_scheme_reset_counters": (C.c_int, [_VP]),
    "hp_scheme_iterate": (C.c_int, [_VP, C.c_uint32]),
    "hp_scheme_prepare_graphs": (C.c_int, [_VP]),
    "hp_scheme_sync": (C.c_int, [_VP]),
    "hp_scheme_read_stats": (C.c_int, [_VP, C.POINTER(HpSchemeStats)]),
    "hp_comm_unique_id": (C.c_int, [_VP]),
    "hp_scheme_attach_comm": (C.c_int, [_VP, _VP, C.c_int, C.c_int]),
    "hp_scheme_peer_export": (C.c_int, [_VP, _VP]),
    "hp_scheme_attach_peers": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "hp_scheme_strip_timing": (C.c_int, [_VP, C.c_int]),
    "hp_scheme_read_strip_phases": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

PEER_BLOB_BYTES = 256
STRIP_PHASES = ("edge_rows", "interior_rows", "halo_wait", "allreduce", "clock")

_lib = None


def prefer_bundled_nccl():
    """Points the library's dlopen of NCCL (HIPIMS_NCCL_LIB, csrc/hp_comm.cpp) at the NCCL that ships with PyTorch.

    One process can hold only one libnccl.so.2: whichever is loaded first serves every later user.  If the library
    loaded the system's (older) NCCL before `import torch`, libtorch_cuda would be bound to it and fail on symbols it
    lacks -- so in a Python process the bundled one is named up front.  The stand-alone C++ host has no such conflict
    and uses the system library."""
    if os.environ.get("HIPIMS_NCCL_LIB"):
        return os.environ["HIPIMS_NCCL_LIB"]
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["HIPIMS_NCCL_LIB"] = cand
                return cand
    except Exception:
        pass
    return None


def load_library():
    """Loads libhipims_cuda.so and declares the ABI.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HipimsCudaError("%s is missing: run `python -m hipims_ocl_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback." % LIB_PATH)
    prefer_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _check(rc):
    if rc < 0:
        raise HipimsCudaError("hipims_cuda error %d: %s" % (rc, load_library().hp_last_error().decode()))
    return rc


def device_count():
    n = C.c_int(0)
    _check(load_library().hp_device_count(C.byref(n)))
    return n.value


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Executor:
    """One CUDA device + stream (replaces CExecutorControlOpenCL + COCLDevice)."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        self.h = _VP()
        _check(self.lib.hp_executor_create(int(device), _VP(stream or 0), C.byref(self.h)))
        name = C.create_string_buffer(256)
        sms, mem = C.c_int(0), C.c_size_t(0)
        _check(self.lib.hp_executor_describe(self.h, name, 256, C.byref(sms), C.byref(mem)))
        self.name, self.sm_count, self.total_mem = name.value.decode(), sms.value, mem.value
        self.device = int(device)
        self._schemes = weakref.WeakSet()   # schemes must be destroyed before their executor

    def finish(self):
        _check(self.lib.hp_executor_finish(self.h))

    def timer_start(self):
        _check(self.lib.hp_executor_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        _check(self.lib.hp_executor_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            for sch in list(self._schemes):
                sch.close()
            self.lib.hp_executor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaScheme:
    """One scheme instance on the device (what a CSchemeGodunov / CSchemeMUSCLHancock /
    CSchemeInertial owns: program, kernels, buffers).  Mirrors oracle.cpu_sim.CpuSim's interface
    so the parity tests drive both with the same code."""

    def __init__(self, executor, cfg, options=0, global_rows=None, row_offset=0, halo_south=0, halo_north=0):
        self.ex, self.cfg, self.lib = executor, cfg, executor.lib
        self.dtype = np.float64 if cfg.precision == "double" else np.float32
        self.rows, self.cols = cfg.rows, cfg.cols
        self.halo_south, self.halo_north = halo_south, halo_north
        c = HpSchemeConfig(struct_size=C.sizeof(HpSchemeConfig), scheme=_SCHEME_ID[cfg.scheme],
                           real_bytes=cfg.real_bytes, quirks=cfg.quirks, options=options,
                           dynamic_timestep=int(cfg.dynamic), friction=int(cfg.friction), cols=cfg.cols, rows=cfg.rows,
                           delta=cfg.delta, courant=cfg.courant, dry_threshold=cfg.dry_threshold,
                           end_time=cfg.end_time, fixed_timestep=cfg.fixed_dt, initial_timestep=cfg.initial_dt,
                           global_rows=cfg.rows if global_rows is None else global_rows, row_offset=row_offset,
                           halo_south=halo_south, halo_north=halo_north)
        self.h = _VP()
        _check(self.lib.hp_scheme_create(executor.h, C.byref(c), C.byref(self.h)))
        executor._schemes.add(self)

    def close(self):
        if self.h:
            if self.ex.h:
                self.lib.hp_scheme_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data ----------------------------------------------------------------------------------
    def upload(self, states, bed, manning):
        s = np.ascontiguousarray(states, dtype=self.dtype).reshape(self.rows, self.cols, 4)
        b = np.ascontiguousarray(bed, dtype=self.dtype).reshape(self.rows, self.cols)
        m = np.ascontiguousarray(manning, dtype=self.dtype).reshape(self.rows, self.cols)
        _check(self.lib.hp_scheme_upload_cells(self.h, _ptr(s), _ptr(b), _ptr(m)))
        self.sync()

    def upload_ptrs(self, states_ptr, bed_ptr, manning_ptr):
        """Raw host pointers (e.g. pinned torch tensors); asynchronous."""
        _check(self.lib.hp_scheme_upload_cells(self.h, _VP(states_ptr), _VP(bed_ptr), _VP(manning_ptr)))

    def download(self, out=None):
        out = np.empty((self.rows, self.cols, 4), dtype=self.dtype) if out is None else out
        _check(self.lib.hp_scheme_download_cells(self.h, _ptr(out)))
        return out

    def download_ptr(self, states_ptr):
        _check(self.lib.hp_scheme_download_cells(self.h, _VP(states_ptr)))

    def download_both(self):
        a = np.empty((self.rows, self.cols, 4), dtype=self.dtype)
        b = np.empty_like(a)
        _check(self.lib.hp_scheme_download_both(self.h, _ptr(a), _ptr(b)))
        return a, b

    def read_rows(self, first_row, count):
        out = np.empty((count, self.cols, 4), dtype=self.dtype)
        _check(self.lib.hp_scheme_read_rows(self.h, first_row, count, _ptr(out)))
        return out

    def derive_raster(self, value, nodata=-9999.0):
        """One output raster (hp.RASTER_* code or the reference's XML name) of the owned rows, NORTH row first, float64."""
        code = hc.RASTER_VALUES[value] if isinstance(value, str) else int(value)
        out = np.empty((self.rows - self.halo_south - self.halo_north, self.cols), dtype=np.float64)
        _check(self.lib.hp_scheme_derive_raster(self.h, code, float(nodata), _ptr(out)))
        return out

    def write_rows(self, first_row, states):
        s = np.ascontiguousarray(states, dtype=self.dtype)
        _check(self.lib.hp_scheme_write_rows(self.h, first_row, s.shape[0], _ptr(s)))
        self.sync()

    # -- clock ---------------------------------------------------------------------------------
    def set_target(self, t):
        _check(self.lib.hp_scheme_set_target_time(self.h, float(t)))

    def set_clock(self, time, timestep, time_hydro=0.0):
        _check(self.lib.hp_scheme_set_clock(self.h, float(time), float(timestep), float(time_hydro)))

    def force_timestep(self, dt):
        _check(self.lib.hp_scheme_force_timestep(self.h, float(dt)))

    def update_timestep(self):
        _check(self.lib.hp_scheme_update_timestep(self.h))

    def reset_counters(self):
        _check(self.lib.hp_scheme_reset_counters(self.h))

    def raw_stats(self):
        st = HpSchemeStats()
        _check(self.lib.hp_scheme_read_stats(self.h, C.byref(st)))
        return st

    def stats(self):
        st = self.raw_stats()
        return {"time": st.time, "timestep": st.timestep, "time_hydro": st.time_hydrological,
                "time_target": st.time_target, "batch_timesteps": st.batch_timesteps,
                "batch_successful": st.batch_successful, "batch_skipped": st.batch_skipped,
                "use_alternate": st.use_alternate}

    # -- boundaries (CBoundary*::prepareBoundary) -------------------------------------------------
    def add_uniform(self, definition, times, values):
        tv = np.ascontiguousarray(np.stack([times, values], axis=1), dtype=np.float64)
        conf = HpBdyUniform(entries=len(times), definition=int(definition), interval=float(times[1] - times[0]),
                            length=float(times[-1]))
        return _check(self.lib.hp_boundary_add_uniform(self.h, C.byref(conf), _ptr(tv)))

    def add_gridded(self, definition, interval, resolution, offset_x, offset_y, frames):
        fr = np.ascontiguousarray(frames, dtype=np.float64)
        conf = HpBdyGridded(interval=interval, resolution=resolution, offset_x=offset_x, offset_y=offset_y,
                            entries=fr.shape[0], definition=int(definition), rows=fr.shape[1], cols=fr.shape[2])
        return _check(self.lib.hp_boundary_add_gridded(self.h, C.byref(conf), _ptr(fr)))

    def add_cell(self, def_depth, def_discharge, cell_ids, series_tdxy):
        rel = np.ascontiguousarray(cell_ids, dtype=np.uint64)
        ts = np.ascontiguousarray(series_tdxy, dtype=np.float64)
        conf = HpBdyCell(entries=ts.shape[0], interval=float(ts[1, 0] - ts[0, 0]), length=float(ts[-1, 0]),
                         relations=len(rel), def_depth=int(def_depth), def_discharge=int(def_discharge))
        return _check(self.lib.hp_boundary_add_cell(self.h, C.byref(conf), _ptr(rel), _ptr(ts)))

    # -- the hot loop --------------------------------------------------------------------------
    def iterate(self, n=1, sync=True):
        _check(self.lib.hp_scheme_iterate(self.h, int(n)))
        if sync:
            self.sync()

    def prepare_graphs(self):
        """Builds the CUDA graphs iterate() replays (collective when a communicator is attached)."""
        _check(self.lib.hp_scheme_prepare_graphs(self.h))

    def sync(self):
        _check(self.lib.hp_scheme_sync(self.h))

    # -- multi-GPU -----------------------------------------------------------------------------
    def attach_comm(self, unique_id, rank, world_size):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(self.lib.hp_scheme_attach_comm(self.h, C.cast(buf, _VP), int(rank), int(world_size)))


    def peer_export(self):
        """This strip's buffers and mailbox as an opaque blob for its peers (hp_scheme_peer_export)."""
        buf = (C.c_char * PEER_BLOB_BYTES)()
        _check(self.lib.hp_scheme_peer_export(self.h, C.cast(buf, _VP)))
        return bytes(buf)

    def attach_peers(self, rank, world_size, blobs):
        """Row strips over peer memory instead of NCCL: `blobs` = peer_export() of every strip, by rank.  A rendezvous --
        every strip must call it, after its upload."""
        assert len(blobs) == world_size and all(len(b) == PEER_BLOB_BYTES for b in blobs)
        buf = (C.c_char * (PEER_BLOB_BYTES * world_size)).from_buffer_copy(b"".join(blobs))
        _check(self.lib.hp_scheme_attach_peers(self.h, int(rank), int(world_size), C.cast(buf, _VP)))

    def strip_timing(self, enable=True):
        _check(self.lib.hp_scheme_strip_timing(self.h, int(bool(enable))))

    def strip_phases(self):
        """{phase: ms per iteration} accumulated since strip_timing(True), and the iteration count."""
        ms = (C.c_double * len(STRIP_PHASES))()
        n = C.c_uint64(0)
        _check(self.lib.hp_scheme_read_strip_phases(self.h, ms, C.byref(n)))
        return dict(zip(STRIP_PHASES, [float(v) for v in ms])), int(n.value)


def comm_unique_id():
    buf = (C.c_char * 128)()
    _check(load_library().hp_comm_unique_id(C.cast(buf, _VP)))
    return bytes(buf)
