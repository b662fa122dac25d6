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


def uniform_splits(grid):
    """Comm::set_proc_grid default: xsplit[i] = i / procgrid (comm.cpp), last = 1.0"""
    return [np.array([i * 1.0 / g for i in range(g)] + [1.0]) for g in grid]


def balance_shift(x, boxlo, boxhi, grid, dims="xyz", niter=10, thresh=1.1, reduce=None, natoms=None):
    """the `balance thresh shift dims Niter stopthresh` command: Balance::shift (src/balance.cpp:632-790) with its recursive
    multisection of the cuts (adjust :832-884, static form rho = 0: every cut moves to the midpoint of its bracket) on the
    lamda coordinates of the atoms; tally (:798-815) counts the atoms per slice -- `reduce` sums the counts over the ranks when
    every rank holds only its own atoms (MPI_Allreduce there).  Returns [xsplit, ysplit, zsplit] (fractions of the box, as
    Comm::xsplit) starting from the uniform cuts."""
    splits = uniform_splits(grid)
    x = np.asarray(x, np.float64)
    tot = lambda a: a if reduce is None else reduce(a)
    if natoms is None:
        natoms = int(tot(np.array([len(x)], np.int64))[0])
    if natoms == 0:
        return splits
    delta = thresh ** (1.0 / len(dims)) - 1.0
    for ch in dims:
        d = "xyz".index(ch)
        n = grid[d]
        if n == 1:
            continue
        lam = (x[:, d] - boxlo[d]) / (boxhi[d] - boxlo[d])
        split = splits[d]

        def tally():
            # Balance::binary: slice i holds split[i] <= value < split[i+1]; below the first / at or above the last cut -> end slices
            idx = np.clip(np.searchsorted(split[:n], lam, side="right") - 1, 0, n - 1)
            count = tot(np.bincount(idx, minlength=n).astype(np.int64))
            return np.concatenate([[0], np.cumsum(count)])
        sums = tally()
        target = np.array([int(1.0 * natoms / n * i + 0.5) for i in range(n)] + [natoms], np.int64)
        lo = np.zeros(n + 1); hi = np.ones(n + 1); losum = np.zeros(n + 1, np.int64); hisum = np.full(n + 1, natoms, np.int64)
        for i in range(1, n):
            for j in range(i, -1, -1):
                if sums[j] <= target[i]:
                    lo[i], losum[i] = split[j], sums[j]
                    break
            for j in range(i, n + 1):
                if sums[j] >= target[i]:
                    hi[i], hisum[i] = split[j], sums[j]
                    break
        for _ in range(niter):
            for i in range(1, n):                      # adjust(): tighten the brackets with the current cuts ...
                if sums[i] <= target[i]:
                    lo[i], losum[i] = split[i], sums[i]
                if sums[i] >= target[i]:
                    hi[i], hisum[i] = split[i], sums[i]
            for i in range(1, n):
                if lo[i] < lo[i - 1]:
                    lo[i], losum[i] = lo[i - 1], losum[i - 1]
            for i in range(n - 1, 0, -1):
                if hi[i] > hi[i + 1]:
                    hi[i], hisum[i] = hi[i + 1], hisum[i + 1]
            change = False
            for i in range(1, n):                      # ... and bisect
                if sums[i] != target[i]:
                    change = True
                    split[i] = 0.5 * (lo[i] + hi[i])
            sums = tally()
            if not change:
                break
            if all(abs(1.0 * (sums[i] - target[i])) / target[i] <= delta for i in range(1, n)):
                break
        if any(split[i] >= split[i + 1] for i in range(n)):
            raise RuntimeError("balance shift: zero-width sub-domain (Balance::shift 'Bad split')")
    return splits


class Brick:
    """one rank's place in the decomposition; splits = [xsplit, ysplit, zsplit] (Comm::xsplit ..., fractions of the box) for
    non-uniform bricks, e.g. from balance_shift()"""

    def __init__(self, world, rank, boxlo, boxhi, dim=3, grid=None, splits=None):
        self.world, self.rank = world, rank
        self.grid = proc_grid(world, boxlo, boxhi, dim, grid)
        self.splits = splits
        px, py, pz = self.grid
        self.myloc = (rank // (py * pz), (rank // pz) % py, rank % pz)          # MPI_Cart_create ordering
        self.procneigh = []
        for d in range(3):
            for step in (-1, +1):
                loc = list(self.myloc)
                loc[d] = (loc[d] + step) % self.grid[d]
                self.procneigh.append(loc[0] * py * pz + loc[1] * pz + loc[2])
        self.sublo, self.subhi = [], []
        for d in range(3):                  # Domain::set_local_box, domain.cpp:306-330
            prd = boxhi[d] - boxlo[d]
            lo_f = self.myloc[d] * 1.0 / self.grid[d] if splits is None else float(splits[d][self.myloc[d]])
            hi_f = (self.myloc[d] + 1) * 1.0 / self.grid[d] if splits is None else float(splits[d][self.myloc[d] + 1])
            self.sublo.append(boxlo[d] + prd * lo_f)
            self.subhi.append(boxlo[d] + prd * hi_f if self.myloc[d] < self.grid[d] - 1 else boxhi[d])

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
