"""Multi-GPU host plumbing: LAMMPS-style brick decomposition, one process per GPU.

Mirrors what the reference computes on the host before any atom moves:
  ProcMap::onelevel_grid   src/procmap.cpp  (processor grid with the smallest sub-domain surface)
  Comm::set_proc_grid      src/comm.cpp     (myloc, procneigh via the Cartesian map)
  Domain::set_local_box    src/domain.cpp   (sublo/subhi = boxlo + prd * xsplit[myloc])
The data path (halo exchange, reverse accumulation, migration) is inside libb200sph.so over
NCCL; torch.distributed is used only to broadcast the NCCL id and to gather results.
"""
import ctypes as C
import numpy as np


def factor3(n, dim=3):
    out = []
    for px in range(1, n + 1):
        if n % px:
            continue
        for py in range(1, n // px + 1):
            if (n // px) % py:
                continue
            pz = n // px // py
            if dim == 2 and pz != 1:
                continue
            out.append((px, py, pz))
    return out


def proc_grid(world, boxlo, boxhi, dim=3, user=None):
    """ProcMap::onelevel_grid + best_factors: minimise the surface area of a sub-domain"""
    if user is not None:
        assert user[0] * user[1] * user[2] == world
        return tuple(user)
    prd = [boxhi[d] - boxlo[d] for d in range(3)]
    area = [prd[0] * prd[1], prd[0] * prd[2], prd[1] * prd[2]]
    best, bestsurf = None, None
    for px, py, pz in factor3(world, dim):
        surf = area[0] / px / py + area[1] / px / pz + area[2] / py / pz
        if bestsurf is None or surf < bestsurf:
            best, bestsurf = (px, py, pz), surf
    return best


class Brick:
    """one rank's place in the decomposition"""

    def __init__(self, world, rank, boxlo, boxhi, dim=3, grid=None):
        self.world, self.rank = world, rank
        self.grid = proc_grid(world, boxlo, boxhi, dim, grid)
        px, py, pz = self.grid
        self.myloc = (rank // (py * pz), (rank // pz) % py, rank % pz)          # MPI_Cart_create ordering
        self.procneigh = []
        for d in range(3):
            for step in (-1, +1):
                loc = list(self.myloc)
                loc[d] = (loc[d] + step) % self.grid[d]
                self.procneigh.append(loc[0] * py * pz + loc[1] * pz + loc[2])
        self.sublo, self.subhi = [], []
        for d in range(3):
            prd = boxhi[d] - boxlo[d]
            self.sublo.append(boxlo[d] + prd * (self.myloc[d] * 1.0 / self.grid[d]))
            self.subhi.append(boxlo[d] + prd * ((self.myloc[d] + 1) * 1.0 / self.grid[d]) if self.myloc[d] < self.grid[d] - 1 else boxhi[d])

    def owns(self, x):
        """atoms of this sub-domain: sublo <= x < subhi (create_atoms / comm_brick.cpp:629)"""
        m = np.ones(len(x), bool)
        for d in range(3):
            m &= (x[:, d] >= self.sublo[d]) & (x[:, d] < self.subhi[d])
        return m


def nccl_id(api, dist=None, src=0):
    """rank `src` creates the NCCL unique id, everybody receives it (torch.distributed object broadcast)"""
    buf = C.create_string_buffer(128)
    if dist is None or dist.get_rank() == src:
        api.check(api.comm_unique_id(buf))
    if dist is None:
        return buf.raw
    box = [buf.raw]
    dist.broadcast_object_list(box, src=src)
    return box[0]
