"""Row strips on several GPUs against ONE GPU, bit for bit, through the CUDA / NCCL path (the reference's
CDomainLink exchange, src/Domain/Links/CDomainLink.cpp:168-270, re-targeted to NCCL halo rows + dt all-reduce).

Each case launches tools/multigpu_check.py under torchrun with one rank per GPU: every rank steps its strip through
the C ABI, rank 0 also runs the whole domain on its own GPU and compares states and clocks for equality.
Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_check(world, *argv):
    port = 29600 + (os.getpid() + hash(argv)) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py")] + [str(a) for a in argv]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out = "\n".join(l for l in (res.stdout + res.stderr).splitlines()
                    if l.strip() and not l.startswith("*****") and "OMP_NUM_THREADS" not in l)
    verdict = [l for l in out.splitlines() if l.startswith("multigpu_check")]
    assert res.returncode == 0, "\n".join(verdict) + "\n" + out[-2000:]
    assert "state IDENTICAL" in out and "clocks IDENTICAL" in out, "\n".join(verdict) + "\n" + out[-2000:]
    return out


# scheme, precision, rows, cols, iterations, boundary set, options
CASES = [
    ("godunov", "double", 256, 192, 60, "cells", 0),
    ("muscl-hancock", "double", 256, 192, 60, "cells", 0),
    ("inertial", "double", 256, 192, 60, "cells", 0),
    ("muscl-hancock", "single", 250, 200, 40, "cells", 0),            # rows not divisible by the rank count
    ("godunov", "double", 256, 192, 41, "rain", 2),                   # direct launches (HP_OPT_NO_GRAPH), odd count
    ("muscl-hancock", "double", 256, 192, 40, "cells", 32),           # split path forced on a small strip
    ("muscl-hancock", "double", 256, 192, 40, "cells", 128),          # two columns per lane (HP_OPT_WIDE_MARCH)
    ("inertial", "single", 250, 200, 40, "cells", 32),                # wide kernel, split path, rows not divisible
    ("inertial", "double", 256, 192, 40, "cells", 64),                # one column per lane (HP_OPT_NARROW_MARCH)
]


@pytest.mark.parametrize("scheme,precision,rows,cols,iters,bdy,options", CASES)
def test_two_strips_equal_one_gpu(scheme, precision, rows, cols, iters, bdy, options):
    if gpu_count() < 2:
        pytest.skip("needs two GPUs")
    run_check(2, scheme, precision, rows, cols, iters, bdy, options)


# the same strips exchanging over PEER MEMORY (hp_scheme_attach_peers: one kernel stores the edge rows into the neighbours'
# halo rows, exchanges the wave-speed maximum through mailboxes and runs the time controller; no NCCL in the loop)
PEER_CASES = [
    ("godunov", "double", 256, 192, 60, "cells", 0),
    ("muscl-hancock", "double", 256, 192, 60, "cells", 0),
    ("inertial", "single", 250, 200, 41, "cells", 0),                 # rows not divisible, odd count
    ("muscl-hancock", "double", 256, 192, 41, "rain", 2),             # direct launches (HP_OPT_NO_GRAPH)
]


@pytest.mark.parametrize("scheme,precision,rows,cols,iters,bdy,options", PEER_CASES)
def test_two_strips_over_peer_memory_equal_one_gpu(scheme, precision, rows, cols, iters, bdy, options):
    if gpu_count() < 2:
        pytest.skip("needs two GPUs")
    run_check(2, scheme, precision, rows, cols, iters, bdy, options, "peer")


def test_strips_over_peer_memory_with_two_neighbours():
    n = gpu_count()
    if n < 3:
        pytest.skip("needs three GPUs (a strip with two neighbours)")
    world = 4 if n >= 4 else 3
    run_check(world, "muscl-hancock", "double", 64 * world + 1, 192, 40, "cells", 0, "peer")
    if n >= 8:
        run_check(8, "muscl-hancock", "double", 1021, 512, 40, "cells", 0, "peer")     # rows that do not divide by the rank count


def test_decision_at_the_small_strip_threshold(monkeypatch):
    """A decomposition straddling the small/large-strip threshold: an edge strip holds own + 2 rows (small by its own
    count), an inner strip own + 4 (large).  The iteration shape -- and with it the order of NCCL calls -- must be decided
    alike on every rank (hp_scheme_attach_comm), or neighbouring ranks issue different NCCL sequences and hang.  The
    threshold (32 Mi cells in production) is lowered through the HIPIMS_SMALL_STRIP_CELLS test hook so that a small domain
    straddles it: 64 owned rows of 192 columns per rank -> 66 x 192 = 12672 (edge) < 12800 < 68 x 192 = 13056 (inner)."""
    n = gpu_count()
    if n < 3:
        pytest.skip("needs three GPUs (a strip with two neighbours)")
    world = 4 if n >= 4 else 3
    monkeypatch.setenv("HIPIMS_SMALL_STRIP_CELLS", "12800")
    run_check(world, "muscl-hancock", "double", 64 * world, 192, 30, "cells", 0)
    monkeypatch.setenv("HIPIMS_SMALL_STRIP_CELLS", "13100")      # every rank small: graphs with the exchange captured
    run_check(world, "muscl-hancock", "double", 64 * world, 192, 40, "cells", 0)
    monkeypatch.setenv("HIPIMS_SMALL_STRIP_CELLS", "100")        # every rank large: split path, direct launches
    run_check(world, "godunov", "double", 64 * world, 192, 40, "cells", 0)


def test_eight_strips():
    if gpu_count() < 8:
        pytest.skip("needs eight GPUs")
    run_check(8, "muscl-hancock", "double", 1024, 512, 40, "cells", 0)
    run_check(8, "godunov", "double", 1021, 512, 40, "rain", 0)      # rows that do not divide by the rank count
