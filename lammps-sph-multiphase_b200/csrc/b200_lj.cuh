// b200_lj.cuh -- pair_style sph/lj on the row path.
//
// PairSPHLJ::compute (/root/reference/src/USER-SPH/pair_sph_lj.cpp:48-182) is a Lucy-kernel momentum / energy pair loop with the
// Lennard-Jones equation of state of Ree (LJEOS2, :303-333).  It differs from sph/taitwater in one result-relevant way: the long-range
// correction `lrc` is added to the row particle's pressure term INSIDE the neighbor loop (`fi += lrc`, :139), so the force of the
// k-th in-cutoff pair of a half-list row carries k corrections -- the list order is part of the result.  The engine therefore
//   * evaluates a pair only in the row of its half-list owner (the entry's ownership bit, frozen at build time) and hands the
//     partner its share with fp64 atomics (the only style in the library that uses them: the partner's row cannot know its rank);
//   * ranks the row's in-cutoff owned entries in the reference's list order: Neighbor::full_bin walks the stencil bins in
//     (dz, dy, dx) order (neigh_stencil.cpp:434-448) and every bin in ascending local index (bin_atoms fills the linked lists from
//     the back, neighbor.cpp:1911-1950); half_from_full keeps that order (neigh_derive.cpp:83-145).  The sort key of an entry is
//     (stencil rank of bin_j - bin_i, local index of j) from the reference-bin coordinates every particle carries in xt.w; the
//     local index of a ghost is replaced by its order code + the root atom's index (ghost_code, b200_comm.cuh).
// One thread per owned row, O(n^2) in the row length for the ranking: sph/lj is the rarely used EOS variant of SURVEY 8(f2),
// correctness first.  Requires a full-list sub-style in the deck (sph/rhosum in every shipped use), otherwise LAMMPS builds the
// half list directly (half_bin_newton) in another order -- refused by build_plan.
#pragma once
#include "b200_common.cuh"
#include "b200_neigh.cuh"
#include "b200_pair.cuh"

#define LJ_MAXROW 512

struct LjArgs {
  int nlocal, stride, dim, sx, sy, sz;
  const unsigned *list, *far; const int *cnt, *numfar;
  const double4 *xt, *vr; const double *e, *cv; const int *orig, *gimage;
  double4 *fd; double *de;
  const PairTab *tab;
  int *overflow;
};

// Ree's fit of the LJ fluid (pair_sph_lj.cpp:303-333): p / rho^2 and the sound speed from rho, T = e / cv.  Same polynomials,
// written as Horner chains in x = rho beta^(1/4).
__device__ __forceinline__ void lj_eos(double rho, double e, double cv, double &p_over_rhosq, double &c)
{
  const double T = e / cv, beta = 1.0 / T, bs = sqrt(beta), x = rho * sqrt(bs);
  const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
  const double pb = fma(fma(fma(fma(11.195, x, -31.816), x, 35.505), x, -18.698), x, 3.492);     // beta polynomial
  const double ps = fma(fma(fma(fma(9.32, x, -17.076), x, 18.525), x, 13.16), x, 5.369);         // sqrt(beta) polynomial
  const double dA = 3.629 + 7.264 * x - beta * pb - bs * ps + 10.4925 * x2 + 11.46 * x2 * x + 2.176 * x8 * x;
  const double qb = fma(fma(fma(-44.78, x, 95.448), x, -71.01), x, 18.698);
  const double qs = fma(fma(fma(37.28, x, -51.228), x, 37.05), x, 13.16);
  const double d2A = 7.264 + 20.985 * x + beta * qb - bs * qs + 34.38 * x2 + 19.584 * x8;
  p_over_rhosq = T * (1.0 + dA * x) / rho;
  const double csq = T * (1.0 + 2.0 * dA * x + d2A * x2);
  c = csq > 0.0 ? sqrt(csq) : 0.0;
}

__global__ void __launch_bounds__(128) k_force_lj(LjArgs A)
{
  __shared__ PairTab T;
  load_tab(&T, A.tab);
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.nlocal) return;
  const double4 pi = A.xt[row], vi = A.vr[row];
  const unsigned long long wi = (unsigned long long)__double_as_longlong(pi.w);
  const int ti = tw_type(wi);
  const int c = A.cnt[row], n_in = c & 0xffff, n_out = c >> 16, n_far = A.numfar[row], ntot = n_in + n_out + n_far;
  const unsigned *p = row_base(A.list, row, A.stride), *pf = row_base(A.far, row, A.stride);
  auto entry = [&](int k) {
    if (k < n_in) return p[(size_t)k * 32];
    if (k < n_in + n_out) return p[(size_t)(A.stride - 1 - (k - n_in)) * 32];
    return pf[(size_t)(k - n_in - n_out) * 32];
  };
  // the row's owned, in-cutoff entries with their place in the reference's list order
  unsigned long long key[LJ_MAXROW]; int jdx[LJ_MAXROW];
  int n = 0;
  const int wy = 2 * A.sy + 1, wx = 2 * A.sx + 1;
  for (int k = 0; k < ntot; k++) {
    const unsigned ent = entry(k);
    if (!(ent & NBR_OWNER_BIT)) continue;
    const int j = ent & NBR_INDEX_MASK, tj = (ent >> NBR_TYPE_SHIFT) & 7;
    const double4 pj = A.xt[j];
    const double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
    if (!(rsq < T.cutsq[ti * MAXT1 + tj])) continue;
    const unsigned long long wj = (unsigned long long)__double_as_longlong(pj.w);
    const int sb = ((tw_bz(wj) - tw_bz(wi) + A.sz) * wy + (tw_by(wj) - tw_by(wi) + A.sy)) * wx + (tw_bx(wj) - tw_bx(wi) + A.sx);
    // order of the atoms of one bin = ascending LAMMPS local index: owned atoms by orig, then the ghosts in the reference's border
    // order, i.e. by (order code, local index of the root atom on its rank) -- see ghost_code(), b200_comm.cuh
    const unsigned long long lj = j < A.nlocal ? (unsigned long long)(unsigned)A.orig[j]
                                                : (((unsigned long long)((A.gimage[j] >> 8) & 0x1ff) << 28) | (unsigned)A.orig[j]);
    if (n == LJ_MAXROW) { atomicExch(A.overflow, 1); return; }
    key[n] = ((unsigned long long)(unsigned)sb << 40) | lj; jdx[n] = j | (tj << NBR_TYPE_SHIFT);
    n++;
  }
  double fi0, ci;
  lj_eos(vi.w, A.e[row], A.cv[row], fi0, ci);
  const double mi = T.mass[ti];
  double fx = 0, fy = 0, fz = 0, adrho = 0, ade = 0;
  for (int a = 0; a < n; a++) {
    const int j = jdx[a] & NBR_INDEX_MASK, tj = (jdx[a] >> NBR_TYPE_SHIFT) & 7, ij = ti * MAXT1 + tj;
    // corrections piled up by the entries the reference visited before this one (:137-139), then this pair's own
    double fi = fi0;
    const double hh = T.h[ij], ih3 = 1.0 / (hh * hh * hh);
    const double lrc = -11.1701 * (ih3 * ih3 * ih3 - 1.5 * ih3);
    for (int b = 0; b < n; b++)
      if (key[b] < key[a]) {
        const double hb = T.h[ti * MAXT1 + ((jdx[b] >> NBR_TYPE_SHIFT) & 7)], ib3 = 1.0 / (hb * hb * hb);
        fi += -11.1701 * (ib3 * ib3 * ib3 - 1.5 * ib3);
      }
    fi += lrc;
    const double4 pj = A.xt[j], vj = A.vr[j];
    double fj, cj;
    lj_eos(vj.w, A.e[j], A.cv[j], fj, cj);
    fj += lrc;
    const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    const double rsq = rsq_nofma(dx, dy, dz);
    double wfd = hh - sqrt(rsq); wfd = T.c0[ij] * wfd * wfd;                        // Lucy (dW/dr)/r (:107-118), constant folded in fill_tab
    const double dvdr = dx * (vi.x - vj.x) + dy * (vi.y - vj.y) + dz * (vi.z - vj.z);
    double fvisc = 0.0;
    if (dvdr < 0.0) fvisc = -T.visc[ij] * (ci + cj) * (hh * dvdr / (rsq + 0.01 * hh * hh)) / (vi.w + vj.w);      // Monaghan 1992 (:146-151)
    const double mj = T.mass[tj];
    const double fpair = -mi * mj * (fi + fj + fvisc) * wfd, dE = -0.5 * fpair * dvdr;
    fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
    adrho += mj * dvdr * wfd; ade += dE;
    double *fj4 = (double *)(A.fd + j);
    atomicAdd(fj4, -dx * fpair); atomicAdd(fj4 + 1, -dy * fpair); atomicAdd(fj4 + 2, -dz * fpair); atomicAdd(fj4 + 3, mi * dvdr * wfd);
    atomicAdd(A.de + j, dE);
  }
  double *f4 = (double *)(A.fd + row);
  atomicAdd(f4, fx); atomicAdd(f4 + 1, fy); atomicAdd(f4 + 2, fz); atomicAdd(f4 + 3, adrho);
  atomicAdd(A.de + row, ade);
}
