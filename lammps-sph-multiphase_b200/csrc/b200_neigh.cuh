// b200_neigh.cuh -- binning, ghost (halo) construction and the neighbor build.
//
// Replaces Neighbor::bin_atoms / full_bin (src/neighbor.cpp:1911-1990,
// src/neigh_full.cpp:241-340), half_from_full_newton (src/neigh_derive.cpp:83-145)
// and CommBrick::borders / forward_comm for periodic self-images
// (src/comm_brick.cpp:444-506,696-864).  All paths relative to /root/reference/.
//
// Layout: owned particles [0,nlocal) sorted by engine cell (cell edge >= cutneighmax,
// 3x3x3 stencil), ghosts [nlocal,nall) sorted by cell as well; inside a cell the
// order is ascending tag (ghosts: tag*32+image), so every summation order in the
// engine is a pure function of the particle set -- no atomics decide an order.
#pragma once
#include "b200_common.cuh"

// ---------------------------------------------------------------- helpers ---
__device__ __forceinline__ double rsq_nofma(double dx, double dy, double dz)
{ // delx*delx + dely*dely + delz*delz exactly as the reference's x86-64 build rounds it
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Neighbor::coord2bin, one dimension (neighbor.cpp:1961-1990)
__device__ __forceinline__ int refbin1(double x, double lo, double hi, double inv, int nbin)
{
  if (x >= hi) return __double2int_rz(__dmul_rn(__dsub_rn(x, hi), inv)) + nbin;
  if (x >= lo) return imin(__double2int_rz(__dmul_rn(__dsub_rn(x, lo), inv)), nbin - 1);
  return __double2int_rz(__dmul_rn(__dsub_rn(x, lo), inv)) - 1;
}
__device__ __forceinline__ unsigned long long make_tw(const Geom &g, int type, double x, double y, double z)
{
  return pack_tw(type, refbin1(x, g.boxlo[0], g.boxhi[0], g.bininv[0], g.nbin[0]),
                 refbin1(y, g.boxlo[1], g.boxhi[1], g.bininv[1], g.nbin[1]),
                 refbin1(z, g.boxlo[2], g.boxhi[2], g.bininv[2], g.nbin[2]));
}
__device__ __forceinline__ int cell1(double x, double lo, double inv, int n)
{
  int c = __double2int_rd((x - lo) * inv);
  return imin(imax(c, 0), n - 1);
}
__device__ __forceinline__ int cell_of(const Geom &g, double x, double y, double z)
{
  int cx = cell1(x, g.clo[0], g.cinv[0], g.nc[0]), cy = cell1(y, g.clo[1], g.cinv[1], g.nc[1]),
      cz = cell1(z, g.clo[2], g.cinv[2], g.nc[2]);
  return (cz * g.nc[1] + cy) * g.nc[0] + cx;
}

// ------------------------------------------------------------------ scan ----
// exclusive scan of int arrays (counts -> offsets); data[n] receives the total.
#define SCAN_T 256
#define SCAN_E 4
__global__ void k_scan_block(int *data, int n, int *sums)
{
  __shared__ int sh[SCAN_T];
  int base = blockIdx.x * SCAN_T * SCAN_E + threadIdx.x * SCAN_E;
  int v[SCAN_E], t = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) { v[k] = (base + k < n) ? data[base + k] : 0; t += v[k]; }
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 1; o < SCAN_T; o <<= 1) {
    int a = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += a;
    __syncthreads();
  }
  int excl = sh[threadIdx.x] - t;
  if (threadIdx.x == SCAN_T - 1) sums[blockIdx.x] = sh[threadIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) { if (base + k < n) data[base + k] = excl; excl += v[k]; }
}
// the per-block totals of k_scan_block -> exclusive offsets in place, grand total to *total: ONE CTA walks them (<= 62 500 totals for
// 64 M elements), which replaces the recursion of round 1 (two more scan levels + three device copies per scan: the ghost construction
// of a 1 M-particle box was 85 launches, profiles/r02_launches_c3.csv)
__global__ void __launch_bounds__(1024) k_scan_sums(int *sums, int nb, int *total)
{
  __shared__ int wsum[32];
  const int tid = threadIdx.x, per = (nb + 1023) / 1024, lo = min(tid * per, nb), hi = min(lo + per, nb);
  int t = 0;
  for (int k = lo; k < hi; k++) t += sums[k];
  int incl = t;                                                   // inclusive scan of the threads' partial sums
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULLMASK, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = wsum[tid], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULLMASK, wi, o); if (tid >= o) wi += v; }
    wsum[tid] = wi - w;                                           // exclusive offset of warp tid
    if (tid == 31) *total = wi;
  }
  __syncthreads();
  int run = wsum[tid >> 5] + incl - t;
  for (int k = lo; k < hi; k++) { int v = sums[k]; sums[k] = run; run += v; }
}
__global__ void k_scan_add(int *data, int n, const int *sums)
{
  int i = blockIdx.x * SCAN_T * SCAN_E + threadIdx.x;
  int add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_E; k++, i += SCAN_T) if (i < n) data[i] += add;
}

// ------------------------------------------------------- owned: pbc + cells --
// Domain::pbc (src/domain.cpp:476-560) then engine-cell id + histogram
__global__ void k_owned_cells(Geom g, int nlocal, double4 *xt, int *cellid, int *cellcnt, int do_pbc, const int *alive)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (alive && !alive[i]) { cellid[i] = -1; return; }     // slot vacated by atom migration
  double4 p = xt[i];
  if (do_pbc) {
    double c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; d++)
      if (g.periodic[d]) {
        if (c[d] < g.boxlo[d]) c[d] = __dadd_rn(c[d], g.prd[d]);
        if (c[d] >= g.boxhi[d]) { c[d] = __dsub_rn(c[d], g.prd[d]); c[d] = fmax(c[d], g.boxlo[d]); }
      }
    p.x = c[0]; p.y = c[1]; p.z = c[2];
    xt[i] = p;
  }
  int c = cell_of(g, p.x, p.y, p.z);
  cellid[i] = c;
  atomicAdd(&cellcnt[c], 1);
}

// Domain::pbc alone (the multi-rank path wraps before atoms migrate)
__global__ void k_pbc(Geom g, int nlocal, double4 *xt)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  double c[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int d = 0; d < 3; d++)
    if (g.periodic[d]) {
      if (c[d] < g.boxlo[d]) c[d] = __dadd_rn(c[d], g.prd[d]);
      if (c[d] >= g.boxhi[d]) { c[d] = __dsub_rn(c[d], g.prd[d]); c[d] = fmax(c[d], g.boxlo[d]); }
    }
  p.x = c[0]; p.y = c[1]; p.z = c[2];
  xt[i] = p;
}

// extent of the owned atoms for Domain::reset_box (domain.cpp:344-370): max(-x) and max(x) per dimension, reduced on order-preserving
// 64-bit keys (min / max of doubles are exact, so the order of the reduction does not matter)
__device__ __forceinline__ unsigned long long ext_key(double v)
{
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_extent_init(unsigned long long *keys) { if (threadIdx.x < 6) keys[threadIdx.x] = ext_key(-1.0e20); }   // BIG, domain.cpp:39
__global__ void k_extent(int nlocal, const double4 *xt, unsigned long long *keys)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[6];
  if (i < nlocal) { double4 p = xt[i]; v[0] = -p.x; v[1] = p.x; v[2] = -p.y; v[3] = p.y; v[4] = -p.z; v[5] = p.z; }
  else { for (int k = 0; k < 6; k++) v[k] = -1.0e20; }
#pragma unroll
  for (int k = 0; k < 6; k++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
  }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 6; k++) { unsigned long long b = ext_key(v[k]); if (b > *(volatile unsigned long long *)(keys + k)) atomicMax(keys + k, b); }
}
__global__ void k_extent_decode(const unsigned long long *keys, double *out)
{
  if (threadIdx.x < 6) { unsigned long long k = keys[threadIdx.x]; out[threadIdx.x] = __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k)); }
}

// scatter element ids into their cell segment (arbitrary order inside a segment; fixed by k_sort_segments)
__global__ void k_scatter(int n, const int *cellid, const int *cellstart, int *cellfill, int *perm)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cellid[i];
  if (c < 0) return;
  perm[cellstart[c] + atomicAdd(&cellfill[c], 1)] = i;
}

// rank-sort every cell segment by key (unique keys) -> deterministic in-cell order; one warp per cell
__global__ void k_sort_segments(int ncells, const int *cellstart, const int *perm_in, int *perm_out,
                                const unsigned long long *key)
{
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= ncells) return;
  int s = cellstart[w], n = cellstart[w + 1] - s;
  for (int a = lane; a < n; a += 32) {
    int ea = perm_in[s + a];
    unsigned long long ka = key[ea];
    int rank = 0;
    for (int b = 0; b < n; b++) rank += key[perm_in[s + b]] < ka;
    perm_out[s + rank] = ea;
  }
}

// ---- Atom::sort (atom.cpp:1555-1650): the spatial sort that defines LAMMPS' local index order ----
// The engine keeps its own (cell-sorted) device order; what the reference's local order decides -- which atom of a pair is the
// half list's "i" (neigh_derive.cpp:101-110), the order in which fix phase_change draws its random numbers, the order of the
// output arrays -- lives in `orig`.  Atom::sort re-numbers the owned atoms bin by bin (bins of half the neighbor cutoff over the
// sub-domain, setup_sort_bins :1659-1730), atoms of one bin in their previous order: k_sort_bin + scan + k_scatter +
// k_sort_segments (key = previous index) + k_sort_assign reproduce exactly that numbering.
struct SortGeom { double lo[3], inv[3]; int n[3]; };
__global__ void k_sort_bin(SortGeom sg, int nslots, const double4 *xt, const int *alive, const int *orig, int *bin, int *cnt, unsigned long long *key)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots) return;
  if (alive && !alive[i]) { bin[i] = -1; return; }
  double4 p = xt[i];
  int ix = (int)((p.x - sg.lo[0]) * sg.inv[0]), iy = (int)((p.y - sg.lo[1]) * sg.inv[1]), iz = (int)((p.z - sg.lo[2]) * sg.inv[2]);     // static_cast<int>: toward zero
  ix = imin(imax(ix, 0), sg.n[0] - 1); iy = imin(imax(iy, 0), sg.n[1] - 1); iz = imin(imax(iz, 0), sg.n[2] - 1);
  int b = iz * sg.n[1] * sg.n[0] + iy * sg.n[0] + ix;
  bin[i] = b; key[i] = (unsigned long long)(unsigned)orig[i];
  atomicAdd(&cnt[b], 1);
}
__global__ void k_sort_assign(int nslots, const int *ntot, const int *sorted, int *orig)      // *ntot = live atoms (the scan's total)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nslots && p < *ntot) orig[sorted[p]] = p;
}

// gather the owned particles into cell order; xt.w gets type + reference bin of the (wrapped) position
struct OwnedArrays {
  double4 *xt, *vr, *vm, *fd, *cgm;
  double *e, *de, *cv;
  int *tag, *mask, *orig;
};
__global__ void k_permute_owned(Geom g, int nlocal, const int *perm, OwnedArrays a, OwnedArrays b, int multiphase,
                                unsigned long long *key)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nlocal) return;
  int i = perm[s];
  double4 p = a.xt[i];
  int type = tw_type(__double_as_longlong(p.w));
  p.w = __longlong_as_double((long long)make_tw(g, type, p.x, p.y, p.z));
  b.xt[s] = p; b.vr[s] = a.vr[i]; b.vm[s] = a.vm[i]; b.fd[s] = a.fd[i];
  if (multiphase) b.cgm[s] = a.cgm[i];
  b.e[s] = a.e[i]; b.de[s] = a.de[i]; b.cv[s] = a.cv[i];
  b.tag[s] = a.tag[i]; b.mask[s] = a.mask[i]; b.orig[s] = a.orig[i];
}
__global__ void k_owned_keys(int nlocal, const int *tag, unsigned long long *key)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nlocal) key[i] = (unsigned long long)(unsigned)tag[i];
}
// remember positions of the build for Neighbor::check_distance (neighbor.cpp:1428-1441)
__global__ void k_store_xhold(int nlocal, const double4 *xt, double *xhold)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  xhold[3 * i] = p.x; xhold[3 * i + 1] = p.y; xhold[3 * i + 2] = p.z;
}

// ---------------------------------------------------------------- ghosts ----
// Ghosts live behind the owned atoms in swap order (b200_comm.cuh).  For the build they are
// addressed through a cell-ordered index list (gorder), sorted inside a cell by tag*32+image.
__global__ void k_ghost_cells(Geom g, int nlocal, int nghost, const double4 *xt, const int *tag, const int *gimage, int *gcell, int *gcellcnt,
                              unsigned long long *gkey)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nghost) return;
  double4 p = xt[nlocal + s];
  int c = cell_of(g, p.x, p.y, p.z);
  gcell[s] = c;
  gkey[s] = ((unsigned long long)(unsigned)tag[nlocal + s] << 5) | (unsigned)(gimage[nlocal + s] & 31);      // image code only (bits 8+ hold the order code, b200_comm.cuh)
  atomicAdd(&gcellcnt[c], 1);
}

// ------------------------------------------------------------ the build -----
#define BUILD_WARPS 4
#define BUILD_CH 128
struct BuildSmem {
  double x[BUILD_CH], y[BUILD_CH], z[BUILD_CH];
  unsigned long long w[BUILD_CH];
  int j[BUILD_CH], o[BUILD_CH];
};
struct BuildArgs {
  Geom g;
  int nlocal, nghost, stride, ntypes1;
  const double4 *xt;
  const int *orig;
  const int *cso, *csg;          // cell starts: owned / ghost (ghost cell segments index gorder)
  const int *gorder;             // ghost slots in cell order (device index = nlocal + gorder[k])
  const double *cutneighsq;      // [MAXTT]
  const double *prunesq;         // [MAXTT] max over the sub-styles of cutsq(ti,tj): splits a row into an inner and an outer zone
  const double *farsq;           // [MAXTT] (cut + margin)^2: entries beyond go to the far rows (see FAR_MARGIN_FRAC)
  unsigned *far; int *numfar;
  unsigned *nbr;
  int hbn;                       // the deck has no full-list sub-style: pair ownership of half_bin_newton (see half_bin_upper)
  int *numneigh;
  int *maxcount;
};

// half_from_full_newton's rule for an (owned i, ghost j) pair (neigh_derive.cpp:121-134): keep iff j is "above/right" of i
__device__ __forceinline__ bool ghost_above(double xi, double yi, double zi, double xj, double yj, double zj)
{
  if (zj < zi) return false;
  if (zj == zi) {
    if (yj < yi) return false;
    if (yj == yi && xj < xi) return false;
  }
  return true;
}

// Neighbor::half_bin_newton (neigh_half_bin.cpp:339-400), the half list of a deck WITHOUT a full-list sub-style: atom i holds the pairs
// with every atom of the bins of the upper half stencil (stencil_half_bin_{2d,3d}_newton, neigh_stencil.cpp:125-158:
// k > 0 || j > 0 || (j == 0 && i > 0)); inside i's own bin the rule is the one of half_from_full_newton (owned atoms behind i in the
// bin's linked list = larger local index, ghosts "above and to the right").  d* = reference bin of j minus reference bin of i.
__device__ __forceinline__ bool half_bin_upper(int dbx, int dby, int dbz)
{
  return dbz > 0 || (dbz == 0 && (dby > 0 || (dby == 0 && dbx > 0)));
}

// One warp per engine cell.  Each lane owns one row particle; the candidates of the
// 3x3x3 stencil are staged through shared memory in chunks and read back as
// broadcasts (one wavefront per 32 pair tests); hits are collected as 32-bit masks.
// Row entries k = 0..n_in-1 (inner zone) and k = stride-1 .. stride-n_out (outer zone); numneigh = n_in | n_out << 16.
// Rows [0,nlocal): Neighbor::full_bin's list of the owned atom (rsq <= cutneighsq, and
// j inside the reference's own bin stencil), each entry tagged with the half-list
// ownership bit frozen at build time.  Rows [nlocal,nall): for a ghost g, the owned
// atoms i whose half list holds (i,g) -- used to accumulate what the reference adds
// to ghost atoms and reverse-communicates.
__global__ void __launch_bounds__(BUILD_WARPS * 32) k_build(BuildArgs A)
{
  __shared__ BuildSmem sm_all[BUILD_WARPS];
  __shared__ double s_cut[MAXTT], s_in[MAXTT], s_far[MAXTT];
  for (int k = threadIdx.x; k < MAXTT; k += blockDim.x) { s_cut[k] = A.cutneighsq[k]; s_in[k] = A.prunesq[k]; s_far[k] = A.farsq[k]; }
  __syncthreads();
  const Geom &g = A.g;
  int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c = blockIdx.x * BUILD_WARPS + wib;
  if (c >= g.ncells) return;
  BuildSmem &sm = sm_all[wib];
  int o0 = A.cso[c], nO = A.cso[c + 1] - o0, g0 = A.csg[c], nG = A.csg[c + 1] - g0;
  if (nO + nG == 0) return;
  int cx = c % g.nc[0], cy = (c / g.nc[0]) % g.nc[1], cz = c / (g.nc[0] * g.nc[1]);
  const double cutmaxsq = g.cutneighmaxsq;
  const double cutsafe = cutmaxsq * (1.0 - 1.0e-9);     // below this, round-off cannot move a pair across a bin-stencil boundary

  for (int pass = 0; pass < 2; pass++) {          // 0: owned rows, 1: ghost rows
    int nrow = pass ? nG : nO, r0 = o0;
    for (int rb = 0; rb < nrow; rb += 32) {
      bool valid = rb + lane < nrow;
      int i = pass ? (valid ? A.nlocal + A.gorder[g0 + rb + lane] : 0) : r0 + rb + lane;
      double xi = 1e300, yi = 1e300, zi = 1e300;
      unsigned long long wi = 0; int oi = 0;
      if (valid) { double4 p = A.xt[i]; xi = p.x; yi = p.y; zi = p.z; wi = (unsigned long long)__double_as_longlong(p.w); if (!pass) oi = A.orig[i]; }
      int ti = tw_type(wi), bxi = tw_bx(wi), byi = tw_by(wi), bzi = tw_bz(wi);
      int cnt = 0, cin = 0, cout = 0, cfar = 0;
      unsigned *frow = A.far + (size_t)(i >> 5) * A.stride * 32 + (i & 31);
      unsigned *row = A.nbr + (size_t)(i >> 5) * A.stride * 32 + (i & 31);   // rows interleaved by 32 (see b200_pair.cuh)
      int fill = 0;

      auto process = [&](int n) {
        // 1. coarse test of the whole chunk (<= 128 candidates, broadcast reads): rsq <= cutneighmax^2 -> 128-bit hit mask
        unsigned long long m0 = 0, m1 = 0;
        {
          int n0 = imin(n, 64);
#pragma unroll 4
          for (int b = 0; b < n0; b++) {
            double rsq = rsq_nofma(xi - sm.x[b], yi - sm.y[b], zi - sm.z[b]);
            m0 |= (unsigned long long)(rsq <= cutmaxsq) << b;
          }
#pragma unroll 4
          for (int b = 64; b < n; b++) {
            double rsq = rsq_nofma(xi - sm.x[b], yi - sm.y[b], zi - sm.z[b]);
            m1 |= (unsigned long long)(rsq <= cutmaxsq) << (b - 64);
          }
        }
        // 2. one flattened loop over this lane's hits (a lane with few hits in the first half moves on to the
        //    second half while others are still busy: the warp runs max-over-lanes iterations, not a sum of maxima)
        while (m0 | m1) {
          int idx;
          if (m0) { idx = __ffsll((long long)m0) - 1; m0 &= m0 - 1; }
          else { idx = 64 + __ffsll((long long)m1) - 1; m1 &= m1 - 1; }
          int j = sm.j[idx];
          if (j == i) continue;
          unsigned long long wj = sm.w[idx];
          double xj = sm.x[idx], yj = sm.y[idx], zj = sm.z[idx];
          double rsq = rsq_nofma(xi - xj, yi - yj, zi - zj);
          int tj = tw_type(wj), tij = ti * MAXT1 + tj;
          if (!(rsq <= s_cut[tij])) continue;
          // j must sit in a bin of the reference's stencil around i's bin (neigh_stencil.cpp:434-448).  A pair that is
          // clearly inside the largest cutoff always does (closest bin distance <= pair distance < cutneighmax); only
          // pairs within rounding distance of cutneighmax can be affected by where round-off put them, so only those
          // are checked against the reference's integer bin coordinates.
          if (rsq >= cutsafe) {
            int dbx = abs(tw_bx(wj) - bxi), dby = abs(tw_by(wj) - byi), dbz = abs(tw_bz(wj) - bzi);
            if (dbx > g.sx || dby > g.sy || dbz > g.sz) continue;
            double ex = dbx ? (dbx - 1) * g.binsize[0] : 0.0, ey = dby ? (dby - 1) * g.binsize[1] : 0.0,
                   ez = dbz ? (dbz - 1) * g.binsize[2] : 0.0;
            if (!(rsq_nofma(ex, ey, ez) < cutmaxsq)) continue;
          }
          unsigned ent = (unsigned)j | ((unsigned)tj << NBR_TYPE_SHIFT);
          const int hbx = tw_bx(wj) - bxi, hby = tw_by(wj) - byi, hbz = tw_bz(wj) - bzi;      // reference bin of j relative to i's
          const bool other_bin = A.hbn && (hbx | hby | hbz);
          if (!pass) {
            bool own = other_bin ? half_bin_upper(hbx, hby, hbz) : ((j < A.nlocal) ? (oi < sm.o[idx]) : ghost_above(xi, yi, zi, xj, yj, zj));
            if (own) ent |= NBR_OWNER_BIT;
          } else {
            // (owned j, ghost i): kept by j's half list?
            if (!(other_bin ? half_bin_upper(-hbx, -hby, -hbz) : ghost_above(xj, yj, zj, xi, yi, zi))) continue;
            ent |= NBR_OWNER_BIT;
          }
          // inner zone (inside the pair cutoff now): filled from the front; outer zone (skin shell, or
          // exactly on the cutoff): filled from the back.  The stage kernels test every entry each step,
          // but hits and misses are clustered, so their warps do not diverge on the heavy pair body.
          // Far rows: entries at least `margin` outside the cutoff.  They cannot come inside before some
          // particle has moved margin/2 since the build, which the integrator tracks (dmaxsq), so the
          // stage kernels skip them -- provably without changing any result -- until that happens.
          if (rsq >= s_far[tij]) { if (cfar < A.stride) frow[(size_t)cfar * 32] = ent; cfar++; }
          else {
            if (cin + cout < A.stride) {
              if (rsq < s_in[tij]) { row[(size_t)cin * 32] = ent; cin++; }
              else { row[(size_t)(A.stride - 1 - cout) * 32] = ent; cout++; }
            }
            cnt++;
          }
        }
      };

      for (int dz = -1; dz <= 1; dz++) {
        int nz = cz + dz; if (nz < 0 || nz >= g.nc[2]) continue;
        for (int dy = -1; dy <= 1; dy++) {
          int ny = cy + dy; if (ny < 0 || ny >= g.nc[1]) continue;
          for (int dx = -1; dx <= 1; dx++) {
            int nx = cx + dx; if (nx < 0 || nx >= g.nc[0]) continue;
            int n = (nz * g.nc[1] + ny) * g.nc[0] + nx;
            for (int rng = 0; rng < (pass ? 1 : 2); rng++) {
              int s = rng ? A.csg[n] : A.cso[n], e = rng ? A.csg[n + 1] : A.cso[n + 1];
              while (s < e) {
                int take = imin(BUILD_CH - fill, e - s);
                for (int k = lane; k < take; k += 32) {
                  int src = rng ? A.nlocal + A.gorder[s + k] : s + k;
                  double4 p = A.xt[src];
                  sm.x[fill + k] = p.x; sm.y[fill + k] = p.y; sm.z[fill + k] = p.z;
                  sm.w[fill + k] = (unsigned long long)__double_as_longlong(p.w);
                  sm.j[fill + k] = src; sm.o[fill + k] = rng ? 0 : A.orig[src];
                }
                fill += take; s += take;
                if (fill == BUILD_CH) { __syncwarp(); process(fill); __syncwarp(); fill = 0; }
              }
            }
          }
        }
      }
      if (fill) { __syncwarp(); process(fill); __syncwarp(); }
      if (valid) { A.numneigh[i] = cin | (cout << 16); A.numfar[i] = cfar; atomicMax(A.maxcount, max(cnt, cfar)); }
    }
  }
}
