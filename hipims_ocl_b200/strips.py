"""Row-strip decomposition (host side).

Replaces the reference's overlapping <domain> stacks and their link zones
(src/Domain/Links/CDomainLink.cpp:286-382, src/Domain/CDomainManager.cpp:427) with a plain
partition of the rows of ONE domain: rank r owns a contiguous block of rows and keeps `halo`
extra rows of each neighbouring strip (1 for Godunov / inertial, 2 for MUSCL-Hancock, whose
corrector needs the neighbours' predictor output).
"""
from dataclasses import dataclass

from . import config as hc


def halo_rows(scheme):
    return 2 if scheme == hc.SCHEME_MUSCL_HANCOCK else 1


@dataclass
class Strip:
    rank: int
    world: int
    global_rows: int
    row_offset: int     # global index of the first owned row
    own_rows: int
    halo_south: int
    halo_north: int

    @property
    def rows(self):      # rows held locally (owned + halo)
        return self.own_rows + self.halo_south + self.halo_north

    @property
    def first_local_row(self):   # global index of local row 0
        return self.row_offset - self.halo_south

    def local_slice(self):
        """Slice of the global row axis this strip holds (incl. halos)."""
        return slice(self.first_local_row, self.first_local_row + self.rows)

    def owned_local_slice(self):
        return slice(self.halo_south, self.halo_south + self.own_rows)

    def owned_global_slice(self):
        return slice(self.row_offset, self.row_offset + self.own_rows)


def make_strip(global_rows, world, rank, scheme):
    """Rows are dealt out as evenly as possible, the first `global_rows % world` strips get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    halo = halo_rows(scheme)
    base, extra = divmod(global_rows, world)
    if base < 2 * halo + 1 and world > 1:
        raise ValueError("strips of %d rows are too thin for a halo of %d" % (base, halo))
    own = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return Strip(rank, world, global_rows, offset, own, halo if rank > 0 else 0, halo if rank < world - 1 else 0)


def split_boundary_cells(cell_ids, cols, strip):
    """Global cell IDs (y * cols + x) that fall inside the rows a strip holds (owned or halo)."""
    lo, hi = strip.first_local_row, strip.first_local_row + strip.rows
    return [c for c in cell_ids if lo <= c // cols < hi]
