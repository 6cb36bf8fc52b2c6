"""Box partition of a structured hex mesh over ranks (host-side set-up, no GPU).

≙ Mesh::CartesianPartitioning + ParMesh + the neighbour tables of GroupCommunicator
(mesh/mesh.cpp:8966-9003, mesh/pmesh.cpp:106, general/communication.hpp:294-298): every rank owns
a box of elements; L-dofs on the closed-box intersections are shared; both sides list them in
ascending global lattice order so that packed buffers line up without any index exchange.
"""
import numpy as np

from . import hex_build

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def rank_coords(rank, grid):
    return (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))


def split(n, parts):
    """element ranges [lo, hi) of `parts` boxes along one direction (as even as possible)"""
    base, rem = divmod(n, parts)
    edges = [0]
    for i in range(parts):
        edges.append(edges[-1] + base + (1 if i < rem else 0))
    return edges


def build_part(GN, grid, rank, p, size=(1.0, 1.0, 1.0), skew=False, want=None):
    """this rank's sub-mesh of the GN[0] x GN[1] x GN[2] global mesh, numbered locally"""
    rc = rank_coords(rank, grid)
    ex = [split(GN[a], grid[a]) for a in range(3)]
    lo = [ex[a][rc[a]] for a in range(3)]
    hi = [ex[a][rc[a] + 1] for a in range(3)]
    n = [hi[a] - lo[a] for a in range(3)]
    kw = {} if want is None else {"want": want}
    m = hex_build(n[0], n[1], n[2], p, *size, skew=skew, part=(*GN, *lo), **kw)
    m["rank_coords"], m["elem_lo"], m["elem_hi"] = rc, lo, hi
    return m


def shared_tables(m, grid, p):
    """(nbr_rank, shared_offsets, shared_ldofs) of b200pa_comm_set_tables for the part `m`"""
    lat = m["lattice"].reshape(-1, 3).astype(np.int64)
    rc = m["rank_coords"]
    lo = [m["elem_lo"][a] * p for a in range(3)]
    hi = [m["elem_hi"][a] * p for a in range(3)]
    PX, PY, PZ = grid
    big = int(lat.max()) + 2
    key = (lat[:, 2] * big + lat[:, 1]) * big + lat[:, 0]
    nbrs = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                q = (rc[0] + dx, rc[1] + dy, rc[2] + dz)
                if not (0 <= q[0] < PX and 0 <= q[1] < PY and 0 <= q[2] < PZ):
                    continue
                sel = np.ones(len(lat), bool)
                for a, d in enumerate((dx, dy, dz)):
                    if d == 1:
                        sel &= lat[:, a] == hi[a]
                    elif d == -1:
                        sel &= lat[:, a] == lo[a]
                idx = np.nonzero(sel)[0]
                idx = idx[np.argsort(key[idx], kind="stable")]
                nbrs.append((q[0] + PX * (q[1] + PY * q[2]), idx.astype(np.int32)))
    nbrs.sort(key=lambda t: t[0])
    ranks = np.array([t[0] for t in nbrs], np.int32)
    offs = np.zeros(len(nbrs) + 1, np.int32)
    for i, t in enumerate(nbrs):
        offs[i + 1] = offs[i] + len(t[1])
    ldofs = np.concatenate([t[1] for t in nbrs]).astype(np.int32) if nbrs else np.zeros(0, np.int32)
    return ranks, offs, ldofs


def global_ids(m, GN, p):
    """global lattice id of every local L-dof (for comparing a partitioned vector with a serial one)"""
    lat = m["lattice"].reshape(-1, 3).astype(np.int64)
    nx, ny = GN[0] * p + 1, GN[1] * p + 1
    return lat[:, 0] + nx * (lat[:, 1] + ny * lat[:, 2])
