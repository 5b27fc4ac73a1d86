"""ctypes front-end to the two CPU checkers (TEST INFRASTRUCTURE ONLY).

`CpuSim("oracle", cfg)` drives oracle/liboracle.so (our restatement, oracle/hipims_oracle.cpp);
`CpuSim("ref", cfg)` drives oracle/_ref/ref_<variant>.so (the reference's own kernel sources
compiled through oracle/ref_shim, see oracle/build_ref.py).  Both expose the C interface of
oracle/hpo_api.h and share the host sequencing of oracle/sim_driver.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (hipims_ocl_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

SCHEMES = {"godunov": 0, "muscl-hancock": 1, "inertial": 2}
_REF_SCHEME = {"godunov": "godunov", "muscl-hancock": "mh", "inertial": "inertial"}

QUIRK_REDUCE_BUFFER_A = 1
QUIRK_BDY_COVERAGE = 2
QUIRK_MH_NO_BOUNDARIES = 4
QUIRK_GODUNOV_DT0_KEEP = 8


class HpoConfig(C.Structure):
    _fields_ = [("cols", C.c_int64), ("rows", C.c_int64), ("delta", C.c_double), ("very_small", C.c_double),
                ("quite_small", C.c_double), ("courant", C.c_double), ("end_time", C.c_double),
                ("fixed_dt", C.c_double), ("initial_dt", C.c_double), ("scheme", C.c_int32), ("dynamic", C.c_int32),
                ("friction", C.c_int32), ("quirks", C.c_uint32), ("threads", C.c_int32), ("real_bytes", C.c_int32)]


class HpoStats(C.Structure):
    _fields_ = [("time", C.c_double), ("timestep", C.c_double), ("time_hydro", C.c_double),
                ("time_target", C.c_double), ("batch_timesteps", C.c_double), ("batch_successful", C.c_uint32),
                ("batch_skipped", C.c_uint32), ("use_alternate", C.c_uint32), ("pad", C.c_uint32)]


class HpoBdyUniform(C.Structure):
    _fields_ = [("entries", C.c_uint32), ("definition", C.c_uint32), ("interval", C.c_double), ("length", C.c_double)]


class HpoBdyGridded(C.Structure):
    _fields_ = [("interval", C.c_double), ("resolution", C.c_double), ("offset_x", C.c_double),
                ("offset_y", C.c_double), ("entries", C.c_uint64), ("definition", C.c_uint64), ("rows", C.c_uint64),
                ("cols", C.c_uint64)]


class HpoBdyCell(C.Structure):
    _fields_ = [("entries", C.c_uint64), ("interval", C.c_double), ("length", C.c_double), ("relations", C.c_uint64),
                ("def_depth", C.c_uint32), ("def_discharge", C.c_uint32)]


def build_oracle(force=False):
    """Compile oracle/hipims_oracle.cpp -> oracle/liboracle.so (g++, a few seconds)."""
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("hipims_oracle.cpp", "sim_driver.h", "hpo_api.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-I", HERE, srcs[0], "-o",
           out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stderr[-4000:])
    return out


def ref_variant(cfg):
    return "%s_%s_%s_%s" % (_REF_SCHEME[cfg.scheme], "f64" if cfg.precision == "double" else "f32",
                            "dyn" if cfg.dynamic else "fix", "fric" if cfg.friction else "nofric")


def ref_library_path(cfg):
    return os.path.join(HERE, "_ref", "ref_%s.so" % ref_variant(cfg))


def ref_available(cfg):
    """True when the matching reference build exists (prebuilt) or can be built here."""
    if os.path.exists(ref_library_path(cfg)):
        return True
    from . import build_ref  # noqa: local import, needs /root/reference
    return build_ref.reference_available()


_LIBS = {}


def _load(path):
    lib = _LIBS.get(path)
    if lib is None:
        lib = C.CDLL(path)
        _LIBS[path] = lib
    return lib


def _bind(lib, prefix):
    """Declare argtypes once per (lib, prefix)."""
    key = "_hpo_bound_" + prefix
    if getattr(lib, key, False):
        return
    vp, cfgp = C.c_void_p, C.POINTER(HpoConfig)
    sig = {
        "create": (vp, [cfgp]), "destroy": (None, [vp]), "upload": (None, [vp, vp, vp, vp]),
        "download": (None, [vp, vp]), "download_both": (None, [vp, vp, vp]), "set_target": (None, [vp, C.c_double]),
        "set_clock": (None, [vp, C.c_double, C.c_double, C.c_double]),
        "add_uniform": (C.c_int, [vp, C.POINTER(HpoBdyUniform), vp]),
        "add_gridded": (C.c_int, [vp, C.POINTER(HpoBdyGridded), vp]),
        "add_cell": (C.c_int, [vp, C.POINTER(HpoBdyCell), vp, vp]),
        "iterate": (None, [vp, C.c_int]), "update_timestep": (None, [vp]), "reset_counters": (None, [vp]),
        "stats": (None, [vp, C.POINTER(HpoStats)]),
        "k_gts": (None, [cfgp, vp, vp, vp, vp, vp]), "k_ine": (None, [cfgp, vp, vp, vp, vp, vp]),
        "k_mch_1st": (None, [cfgp, vp, vp, vp, vp, vp, vp, vp]),
        "k_mch_2nd": (None, [cfgp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "k_reduce": (C.c_double, [cfgp, vp, vp]),
        "k_advance": (None, [cfgp, vp, vp, C.c_double]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype, fn.argtypes = res, args
    setattr(lib, key, True)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CpuSim:
    """One domain on one of the CPU checkers.  `cfg` is a hipims_ocl_b200.SchemeConfig (duck-typed)."""

    def __init__(self, backend, cfg, threads=0):
        self.cfg = cfg
        self.dtype = np.float64 if cfg.precision == "double" else np.float32
        if backend == "oracle":
            self.lib = _load(build_oracle())
            self.prefix = "hpo_f64_" if cfg.precision == "double" else "hpo_f32_"
        elif backend == "ref":
            path = ref_library_path(cfg)
            from . import build_ref
            if build_ref.reference_available():
                build_ref.build_variant(ref_variant(cfg))     # no-op when up to date
            elif not os.path.exists(path):
                raise RuntimeError("reference library %s missing and /root/reference not available" % path)
            self.lib = _load(path)
            self.prefix = "hpo_ref_"
        else:
            raise ValueError(backend)
        self.backend = backend
        _bind(self.lib, self.prefix)
        self.c = HpoConfig(cols=cfg.cols, rows=cfg.rows, delta=cfg.delta, very_small=cfg.dry_threshold,
                           quite_small=cfg.dry_threshold * 10, courant=cfg.courant, end_time=cfg.end_time,
                           fixed_dt=cfg.fixed_dt, initial_dt=cfg.initial_dt, scheme=SCHEMES[cfg.scheme],
                           dynamic=int(cfg.dynamic), friction=int(cfg.friction), quirks=cfg.quirks, threads=threads,
                           real_bytes=8 if cfg.precision == "double" else 4)
        self.h = C.c_void_p(self._f("create")(C.byref(self.c)))
        self._keep = []

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data ----------------------------------------------------------------------------------
    def upload(self, states, bed, manning):
        s = np.ascontiguousarray(states, dtype=self.dtype).reshape(self.cfg.rows, self.cfg.cols, 4)
        b = np.ascontiguousarray(bed, dtype=self.dtype).reshape(self.cfg.rows, self.cfg.cols)
        m = np.ascontiguousarray(manning, dtype=self.dtype).reshape(self.cfg.rows, self.cfg.cols)
        self._f("upload")(self.h, _ptr(s), _ptr(b), _ptr(m))

    def download(self):
        out = np.empty((self.cfg.rows, self.cfg.cols, 4), dtype=self.dtype)
        self._f("download")(self.h, _ptr(out))
        return out

    def download_both(self):
        a = np.empty((self.cfg.rows, self.cfg.cols, 4), dtype=self.dtype)
        b = np.empty_like(a)
        self._f("download_both")(self.h, _ptr(a), _ptr(b))
        return a, b

    # -- clock ---------------------------------------------------------------------------------
    def set_target(self, t):
        self._f("set_target")(self.h, float(t))

    def set_clock(self, time, timestep, time_hydro=0.0):
        self._f("set_clock")(self.h, float(time), float(timestep), float(time_hydro))

    def stats(self):
        st = HpoStats()
        self._f("stats")(self.h, C.byref(st))
        return {k: getattr(st, k) for k, _ in HpoStats._fields_ if k != "pad"}

    # -- boundaries ----------------------------------------------------------------------------
    def add_uniform(self, definition, times, values):
        tv = np.ascontiguousarray(np.stack([times, values], axis=1), dtype=np.float64)
        conf = HpoBdyUniform(entries=len(times), definition=int(definition), interval=float(times[1] - times[0]),
                             length=float(times[-1]))
        return self._f("add_uniform")(self.h, C.byref(conf), _ptr(tv))

    def add_gridded(self, definition, interval, resolution, offset_x, offset_y, frames):
        fr = np.ascontiguousarray(frames, dtype=np.float64)  # [entries, grid_rows, grid_cols]
        conf = HpoBdyGridded(interval=interval, resolution=resolution, offset_x=offset_x, offset_y=offset_y,
                             entries=fr.shape[0], definition=int(definition), rows=fr.shape[1], cols=fr.shape[2])
        return self._f("add_gridded")(self.h, C.byref(conf), _ptr(fr))

    def add_cell(self, def_depth, def_discharge, cell_ids, series_tdxy):
        rel = np.ascontiguousarray(cell_ids, dtype=np.uint64)
        ts = np.ascontiguousarray(series_tdxy, dtype=np.float64)  # [entries, 4] = t, depth|fsl, Qx, Qy
        conf = HpoBdyCell(entries=ts.shape[0], interval=float(ts[1, 0] - ts[0, 0]), length=float(ts[-1, 0]),
                          relations=len(rel), def_depth=int(def_depth), def_discharge=int(def_discharge))
        return self._f("add_cell")(self.h, C.byref(conf), _ptr(rel), _ptr(ts))

    # -- stepping ------------------------------------------------------------------------------
    def iterate(self, n=1):
        self._f("iterate")(self.h, int(n))

    def update_timestep(self):
        self._f("update_timestep")(self.h)

    def reset_counters(self):
        self._f("reset_counters")(self.h)

    # -- single kernels on caller arrays (kernel-by-kernel cross checks) -----------------------
    def k_step(self, dt, bed, src, dst, manning):
        name = "k_gts" if self.cfg.scheme == "godunov" else "k_ine"
        d = np.array([dt], dtype=self.dtype)
        self._f(name)(C.byref(self.c), _ptr(d), _ptr(bed), _ptr(src), _ptr(dst), _ptr(manning))

    def k_step_cached(self, dt, bed, src, dst, manning):
        """The scheme's step kernel in its local-memory form (gts_cacheEnabled / ine_cacheEnabled), run a work-group at a time
        by the shim; only the libraries built from the reference sources have it."""
        fn = getattr(self.lib, "hpo_ref_k_step_cached")
        fn.restype, fn.argtypes = None, [C.POINTER(HpoConfig)] + [C.c_void_p] * 5
        d = np.array([dt], dtype=self.dtype)
        fn(C.byref(self.c), _ptr(d), _ptr(bed), _ptr(src), _ptr(dst), _ptr(manning))

    def k_mch_1st_cached(self, dt, bed, state, faces):
        """mch_1st_cachePrediction, a work-group at a time (reference libraries only)."""
        fn = getattr(self.lib, "hpo_ref_k_mch_1st_cached")
        fn.restype, fn.argtypes = None, [C.POINTER(HpoConfig)] + [C.c_void_p] * 7
        d = np.array([dt], dtype=self.dtype)
        fn(C.byref(self.c), _ptr(d), _ptr(bed), _ptr(state), *[_ptr(f) for f in faces])

    def k_mch_1st(self, dt, bed, state, faces):
        d = np.array([dt], dtype=self.dtype)
        self._f("k_mch_1st")(C.byref(self.c), _ptr(d), _ptr(bed), _ptr(state), *[_ptr(f) for f in faces])

    def k_mch_2nd(self, dt, state, bed, manning, faces):
        d = np.array([dt], dtype=self.dtype)
        self._f("k_mch_2nd")(C.byref(self.c), _ptr(d), _ptr(state), _ptr(bed), _ptr(manning),
                             *[_ptr(f) for f in faces])

    def k_reduce(self, state, bed, rows=None):
        c = self.c
        if rows is not None:   # reduce over a sub-block of whole rows
            c = HpoConfig.from_buffer_copy(self.c)
            c.rows = rows
        return self._f("k_reduce")(C.byref(c), _ptr(state), _ptr(bed))

    def k_advance(self, clock, counters, vmax):
        """clock: 5 reals {time, timestep, time_hydro, target, batch}; counters: 2 uint32."""
        self._f("k_advance")(C.byref(self.c), _ptr(clock), _ptr(counters), float(vmax))
