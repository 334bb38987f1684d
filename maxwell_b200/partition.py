"""Host-side slab partition and halo plan (replaces MxGrid::partition + the Epetra_Import that
Epetra_CrsMatrix::FillComplete builds; reference src/MxGrid.cpp:100-193, MxCrsMatrix.cpp:325-342).

Pure numpy: this is the same plan libmxgpu computes inside mxg_crs_create (mxg_spmv.cu:
planHalo), kept here so callers can split a global operator into per-rank row blocks and so
the plan logic is testable without a GPU (tests/test_partition_gloo.py).
"""
import numpy as np


def slab_cuts(row_gids, n_global, nx_planes, nranks):
    """Row offsets [c_0=0, c_1, ..., c_P=n] of an x-slab partition with balanced DOF counts.

    GID = comp + ncomp*((x*(Ny+1)+y)*(Nz+1)+z) (MxGrid.h:96-114, MxGridField.hpp:142-145), so an
    x-plane is one contiguous GID range of n_global / nx_planes ids and each rank owns a
    contiguous GID range. Cuts fall on plane boundaries.
    """
    row_gids = np.asarray(row_gids)
    plane = n_global // nx_planes
    first_of_plane = np.searchsorted(row_gids, np.arange(nx_planes + 1) * plane)
    cuts = [0]
    for r in range(1, nranks):
        target = r * len(row_gids) / nranks
        x = int(np.argmin(np.abs(first_of_plane - target)))
        cuts.append(max(int(first_of_plane[x]), cuts[-1]))
    cuts.append(len(row_gids))
    return cuts


def local_block(rowptr, col_gids, vals, r0, r1):
    """Rows [r0, r1) of a global CSR (global column ids) as a stand-alone CSR."""
    p0, p1 = int(rowptr[r0]), int(rowptr[r1])
    return (np.ascontiguousarray(rowptr[r0:r1 + 1]) - p0, np.ascontiguousarray(col_gids[p0:p1]),
            np.ascontiguousarray(vals[p0:p1]))


class HaloPlan:
    """Ghost columns of one rank's row block and who owns them.

    ghosts     sorted unique column GIDs not owned by this rank; the first g_lo precede the
               rank's own GID range, the rest follow it
    ext_col    per-entry column index in the extended local space [-g_lo, n_loc + g_hi)
    recv[q]    (start, count) segment of `ghosts` owned by rank q
    """

    def __init__(self, my_gids, col_gids, all_ranges):
        my_gids = np.asarray(my_gids, dtype=np.int64)
        col_gids = np.asarray(col_gids, dtype=np.int64)
        n_loc = len(my_gids)
        pos = np.searchsorted(my_gids, col_gids)
        pos_c = np.minimum(pos, max(n_loc - 1, 0))
        local = (my_gids[pos_c] == col_gids) if n_loc else np.zeros(len(col_gids), dtype=bool)
        self.ghosts = np.unique(col_gids[~local])
        lo = my_gids[0] if n_loc else 0
        self.g_lo = int(np.searchsorted(self.ghosts, lo)) if n_loc else 0
        self.g_hi = len(self.ghosts) - self.g_lo
        gpos = np.searchsorted(self.ghosts, col_gids[~local])
        ext = np.empty(len(col_gids), dtype=np.int64)
        ext[local] = pos[local]
        ext[~local] = np.where(gpos < self.g_lo, gpos - self.g_lo, n_loc + (gpos - self.g_lo))
        self.ext_col = ext
        self.n_loc = n_loc
        self.recv = {}
        for q, (qlo, qhi) in enumerate(all_ranges):
            if qlo > qhi:
                continue
            b = int(np.searchsorted(self.ghosts, qlo, side="left"))
            e = int(np.searchsorted(self.ghosts, qhi, side="right"))
            if e > b:
                self.recv[q] = (b, e - b)
        covered = sum(c for _, c in self.recv.values())
        if covered != len(self.ghosts):
            raise ValueError("%d ghost columns have no owner" % (len(self.ghosts) - covered))

    def send_indices(self, my_gids, requested_gids):
        """Local indices of the GIDs another rank asked for (they must all be owned here)."""
        my_gids = np.asarray(my_gids, dtype=np.int64)
        idx = np.searchsorted(my_gids, requested_gids)
        if np.any(idx >= len(my_gids)) or np.any(my_gids[np.minimum(idx, len(my_gids) - 1)] != requested_gids):
            raise ValueError("asked for a GID this rank does not own")
        return idx

    def extended_x(self, x_local, ghost_values):
        """[lower ghosts | local | upper ghosts] so ext_col + g_lo indexes it directly."""
        return np.concatenate([ghost_values[:self.g_lo], x_local, ghost_values[self.g_lo:]])
