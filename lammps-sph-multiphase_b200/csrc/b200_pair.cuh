// b200_pair.cuh -- the pair-style stages: one warp per neighbor row.
//
// Each kernel restates one (or a fused group of) PairSPH*::compute loops of
// /root/reference/src/USER-SPH (file:line cited at each formula).  A warp owns
// one particle ("row"): lanes stride over the row's neighbor entries (coalesced
// 4-byte reads), gather the neighbor's packed state (32-byte aligned double4
// records -> one sector each), apply the strict `rsq < cutsq` test of the pair
// style, compact the hits with a ballot into a per-warp shared-memory queue and
// process them 32 at a time with full lane utilisation.  Per-lane partial sums
// are combined with a fixed shuffle tree, so no atomics touch f/drho/de and the
// summation order is a pure function of the neighbor row.
//
// The reference visits every pair once (half list, newton on) and updates both
// atoms; here every owned row evaluates its own side of each pair.  All pair
// formulas are bitwise symmetric under i<->j except two quirks, which take the
// frozen half-list ownership bit of the entry into account:
//   - sph/taitwater/multiphase evaluates p_j with gamma[itype]   (pair_sph_taitwater_multiphase.cpp:148)
//   - sph/heatconduction/phasechange clamps Ti then Tj sequentially (..._phasechange.cpp:124-129)
// Pairs with a ghost are evaluated exactly where the reference evaluates them (the
// side whose half list holds the pair, neigh_derive.cpp:121-134) because ghost rho /
// colorgradient can be one step stale in the multiphase styles;
// ghost rows accumulate what the reference adds to ghost atoms before reverse_comm.
#pragma once
#include "b200_common.cuh"
#include "b200_neigh.cuh"

#define PAIR_WARPS 8
#define EPSILON_CG 1.0e-12   // pair_sph_surfacetension.cpp EPSILON

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(FULLMASK, v, o);
  return v;
}

// quintic spline, sph_kernel_quintic.cpp:17-73, argument q = 3 r / h, WITHOUT the norm factor
__device__ __forceinline__ double quintic_w(double q)
{
  double a = 3.0 - q, b = 2.0 - q, c = 1.0 - q;
  double a2 = a * a, b2 = b * b, c2 = c * c;
  double w = a2 * a2 * a;
  if (q < 2.0) w -= 6.0 * (b2 * b2 * b);
  if (q < 1.0) w += 15.0 * (c2 * c2 * c);
  return q < 3.0 ? w : 0.0;
}
__device__ __forceinline__ double quintic_dw(double q)
{
  double q2 = q * q;
  if (q < 1.0) return (-50.0 * q2 + 120.0 * q) * q2 - 120.0 * q;
  if (q < 2.0) return ((25.0 * q - 180.0) * q + 450.0) * q2 - 420.0 * q + 75.0;
  if (q < 3.0) return ((-5.0 * q + 60.0) * q - 270.0) * q2 + 540.0 * q - 405.0;
  return 0.0;
}

// ---- warp-cooperative traversal of one neighbor row --------------------------
// test(entry) -> bool hit (strict cutoff etc.), proc(entry) accumulates one hit.
template <bool COMPACT, class Test, class Proc>
__device__ __forceinline__ void for_each_hit(const unsigned *__restrict__ row, int n, int lane, unsigned *queue, Test test, Proc proc)
{
  if (!COMPACT) {
    for (int k = lane; k < n; k += 32) { unsigned ent = row[k]; if (test(ent)) proc(ent); }
    return;
  }
  int qn = 0;
  unsigned lt = (1u << lane) - 1u;
  for (int k0 = 0; k0 < n; k0 += 32) {
    int k = k0 + lane;
    unsigned ent = 0; bool hit = false;
    if (k < n) { ent = row[k]; hit = test(ent); }
    unsigned bal = __ballot_sync(FULLMASK, hit);
    if (hit) queue[qn + __popc(bal & lt)] = ent;
    qn += __popc(bal);
    __syncwarp();
    if (qn >= 32) {
      proc(queue[lane]);
      __syncwarp();
      unsigned rest = (lane < qn - 32) ? queue[32 + lane] : 0u;
      __syncwarp();
      if (lane < qn - 32) queue[lane] = rest;
      qn -= 32;
      __syncwarp();
    }
  }
  if (lane < qn) proc(queue[lane]);
}

struct PairArgs {
  int nlocal, nall, stride, dim, multiphase;
  const unsigned *nbr;
  const int *numneigh;
  const double4 *xt, *vr, *vm, *cgm, *dq;
  const double *e, *cv;
  double4 *vr_out, *cg_out, *fd;
  double *de;
  const PairTab *tab[4];   // device pointers; [0] density / colorgradient, or the kinds of a fused force pass
};

// ------------------------------------------------------------------ density --
// MP=false: PairSPHRhoSum::compute            pair_sph_rhosum.cpp:112-197   (quadric kernel, per-type mass)
// MP=true : PairSPHRhoSumMultiphase::compute  pair_sph_rhosum_multiphase.cpp:113-168 (quintic, number density * own mass)
template <bool MP>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_rhosum(PairArgs A)
{
  __shared__ PairTab T;
  __shared__ unsigned queue[PAIR_WARPS][64];
  for (int k = threadIdx.x; k < (int)(sizeof(PairTab) / 4); k += blockDim.x) ((int *)&T)[k] = ((const int *)A.tab[0])[k];
  __syncthreads();
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int nw = gridDim.x * PAIR_WARPS;
  for (int i = blockIdx.x * PAIR_WARPS + wib; i < A.nlocal; i += nw) {
    double4 pi = A.xt[i];
    int ti = tw_type(__double_as_longlong(pi.w));
    if (T.iskip[ti]) continue;                       // atoms of skipped types keep their integrated rho
    const double *cutrow = &T.cutsq[ti * MAXT1];
    double acc = 0.0;
    auto test = [&](unsigned ent) {
      double4 pj = A.xt[ent & NBR_INDEX_MASK];
      double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      return rsq < cutrow[tw_type(__double_as_longlong(pj.w))];
    };
    auto proc = [&](unsigned ent) {
      double4 pj = A.xt[ent & NBR_INDEX_MASK];
      int tj = tw_type(__double_as_longlong(pj.w));
      double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      int ij = ti * MAXT1 + tj;
      if (!MP) {
        double wf = 1.0 - rsq * T.c1[ij];            // 1 - r^2/h^2
        wf = wf * wf; wf = wf * wf;
        acc += T.mass[tj] * (T.c0[ij] * wf);         // C_d (1-r^2/h^2)^4 / h^d
      } else {
        acc += T.c0[ij] * quintic_w(3.0 * (sqrt(rsq) * T.c1[ij]));
      }
    };
    for_each_hit<true>(A.nbr + (size_t)i * A.stride, A.numneigh[i], lane, queue[wib], test, proc);
    acc = warp_sum(acc);
    if (lane == 0) {
      double rho;
      if (!MP) rho = T.mass[ti] * T.self0[ti] + acc;
      else rho = (T.self0[ti] + acc) * A.vm[i].w;     // rho[i] *= imass (:170)
      A.vr_out[i].w = rho;
    }
  }
}

// per-particle derived quantities for the gather records (all nall rows, ghosts keep their possibly stale rho/cg):
//  single-phase: q0 = Tait term B((rho/rho0)^7-1)/rho^2 (pair_sph_taitwater.cpp:118-120), q1 = e
//  multiphase  : q0 = pressure (pair_sph_taitwater_multiphase.cpp:289-292), q1 = V^2 = (m/rho)^2, q2 = T = e/cv, q3 = 1/|cg| or 0
__global__ void k_derive(int nall, int multiphase, const PairTab *tait, const double4 *xt, const double4 *vr, const double4 *cgm,
                         const double *e, const double *cv, double4 *dq)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nall) return;
  int t = tw_type(__double_as_longlong(xt[i].w));
  double rho = vr[i].w;
  double4 q = make_double4(0.0, 0.0, 0.0, 0.0);
  if (!multiphase) {
    if (tait) {
      double tmp = rho / tait->rho0[t], fi = tmp * tmp * tmp;
      q.x = tait->B[t] * (fi * fi * tmp - 1.0) / (rho * rho);
    }
    q.y = e[i];
  } else {
    double4 c = cgm[i];
    if (tait) q.x = tait->B[t] * (pow(rho / tait->rho0[t], tait->gamma[t]) - tait->rb[t]);
    double V = c.w / rho;
    q.y = V * V;
    q.z = e[i] / cv[i];
    double a = sqrt(c.x * c.x + c.y * c.y + c.z * c.z);
    q.w = a > EPSILON_CG ? 1.0 / a : 0.0;
  }
  dq[i] = q;
}

// PairSPHColorGradient::compute, pair_sph_colorgradient.cpp:119-184 (full list, owned rows).
// dphi = -W'(r) alpha sigma_i / sigma_j^2, sigma = rho/m  ->  1/sigma_j^2 = V_j^2 (dq.y)
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_colorgradient(PairArgs A)
{
  __shared__ PairTab T;
  __shared__ unsigned queue[PAIR_WARPS][64];
  for (int k = threadIdx.x; k < (int)(sizeof(PairTab) / 4); k += blockDim.x) ((int *)&T)[k] = ((const int *)A.tab[0])[k];
  __syncthreads();
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int nw = gridDim.x * PAIR_WARPS;
  for (int i = blockIdx.x * PAIR_WARPS + wib; i < A.nlocal; i += nw) {
    double4 pi = A.xt[i];
    int ti = tw_type(__double_as_longlong(pi.w));
    if (T.iskip[ti]) continue;
    const double *cutrow = &T.cutsq[ti * MAXT1];
    double ax = 0.0, ay = 0.0, az = 0.0;
    auto test = [&](unsigned ent) {
      double4 pj = A.xt[ent & NBR_INDEX_MASK];
      double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      return rsq < cutrow[tw_type(__double_as_longlong(pj.w))];
    };
    auto proc = [&](unsigned ent) {
      int j = ent & NBR_INDEX_MASK;
      double4 pj = A.xt[j];
      int ij = ti * MAXT1 + tw_type(__double_as_longlong(pj.w));
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      double rsq = rsq_nofma(dx, dy, dz), r = sqrt(rsq), rinv = 1.0 / r;
      double wfd = quintic_dw(3.0 * (r * T.c1[ij])) * T.c0[ij];       // dW/dr
      double s = -wfd * T.visc[ij] * A.dq[j].y * rinv;                // -W' alpha / sigma_j^2 / r   (sigma_i applied at the end)
      ax += s * dx; ay += s * dy; az += s * dz;
    };
    for_each_hit<true>(A.nbr + (size_t)i * A.stride, A.numneigh[i], lane, queue[wib], test, proc);
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane == 0) {
      double sigmai = A.vr[i].w / A.vm[i].w;
      double4 c = A.cg_out[i];
      c.x = ax * sigmai; c.y = ay * sigmai; c.z = (A.dim == 3) ? az * sigmai : 0.0;
      A.cg_out[i] = c;
    }
  }
}

// ------------------------------------------------------------ fused forces ---
// One pass over the rows for every force-type sub-style of a hybrid/overlay group:
//  K_TAIT   PairSPHTaitwater::compute            pair_sph_taitwater.cpp:101-196
//  K_MORRIS PairSPHTaitwaterMorris::compute      pair_sph_taitwater_morris.cpp:102-196
//  K_HEAT   PairSPHHeatConduction::compute       pair_sph_heatconduction.cpp:76-132
//  K_TAITMP PairSPHTaitwaterMultiphase::compute  pair_sph_taitwater_multiphase.cpp:103-182
//  K_SURF   PairSPHSurfaceTension::compute       pair_sph_surfacetension.cpp:81-190
//  K_HEATMP PairSPHHeatConductionMultiPhase      pair_sph_heatconduction_multiphase.cpp:78-127
//  K_HEATPC PairSPHHeatConductionPhaseChange     pair_sph_heatconduction_phasechange.cpp:83-139
// Rows [0,nlocal) are owned particles, rows [nlocal,nall) ghosts (see k_build).
template <int KINDS, bool DIM3>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_force(PairArgs A)
{
  constexpr int NK = ((KINDS & K_TAIT) ? 1 : 0) + ((KINDS & K_MORRIS) ? 1 : 0) + ((KINDS & K_TAITMP) ? 1 : 0) +
                     ((KINDS & K_SURF) ? 1 : 0) + ((KINDS & K_HEAT) ? 1 : 0) + ((KINDS & K_HEATMP) ? 1 : 0) + ((KINDS & K_HEATPC) ? 1 : 0);
  constexpr bool MP = (KINDS & (K_TAITMP | K_SURF | K_HEATMP | K_HEATPC)) != 0;
  constexpr bool NEED_V = (KINDS & (K_TAIT | K_MORRIS | K_TAITMP)) != 0;
  __shared__ PairTab T[NK];
  __shared__ unsigned queue[PAIR_WARPS][64];
  for (int t = 0; t < NK; t++)
    for (int k = threadIdx.x; k < (int)(sizeof(PairTab) / 4); k += blockDim.x) ((int *)&T[t])[k] = ((const int *)A.tab[t])[k];
  __syncthreads();
  // slot of each kind inside T[] (host passes the tables in this canonical order)
  constexpr int I_FLUID = 0;                                                        // TAIT | MORRIS | TAITMP (at most one)
  constexpr int I_SURF = (KINDS & (K_TAIT | K_MORRIS | K_TAITMP)) ? 1 : 0;
  constexpr int I_HEAT = I_SURF + ((KINDS & K_SURF) ? 1 : 0);                        // HEAT | HEATMP | HEATPC (at most one)

  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int nw = gridDim.x * PAIR_WARPS;
  for (int i = blockIdx.x * PAIR_WARPS + wib; i < A.nall; i += nw) {
    const bool ghostrow = i >= A.nlocal;
    double4 pi = A.xt[i];
    int ti = tw_type(__double_as_longlong(pi.w));
    double4 vi = make_double4(0, 0, 0, 0), qi = A.dq[i], ci = make_double4(0, 0, 0, 0);
    vi = A.vr[i];
    if (MP) ci = A.cgm[i];
    double mi = MP ? ci.w : T[0].mass[ti];
    double fx = 0, fy = 0, fz = 0, adrho = 0, ade = 0;

    auto test = [&](unsigned ent) {
      int j = ent & NBR_INDEX_MASK;
      if (!ghostrow && j >= A.nlocal && !(ent & NBR_OWNER_BIT)) return false;   // the other side owns this ghost pair
      double4 pj = A.xt[j];
      int ij = ti * MAXT1 + tw_type(__double_as_longlong(pj.w));
      double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      bool hit = false;
#pragma unroll
      for (int t = 0; t < NK; t++) hit |= rsq < T[t].cutsq[ij];
      return hit;
    };
    auto proc = [&](unsigned ent) {
      int j = ent & NBR_INDEX_MASK;
      // the half-list owner ("i" of the reference loop) of this pair: the row particle or the neighbor
      const bool row_owns = ghostrow ? false : ((ent & NBR_OWNER_BIT) != 0);
      double4 pj = A.xt[j], vj = A.vr[j], qj = A.dq[j];
      int tj = tw_type(__double_as_longlong(pj.w));
      int ij = ti * MAXT1 + tj;
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      double rsq = rsq_nofma(dx, dy, dz);
      double r = sqrt(rsq);
      double4 cj = make_double4(0, 0, 0, 0);
      if (MP) cj = A.cgm[j];
      double mj = MP ? cj.w : T[0].mass[tj];
      double dvx = 0, dvy = 0, dvz = 0, dvdr = 0;
      if (NEED_V) { dvx = vi.x - vj.x; dvy = vi.y - vj.y; dvz = vi.z - vj.z; dvdr = dx * dvx + dy * dvy + dz * dvz; }
      double rhoi = vi.w, rhoj = vj.w;

      if ((KINDS & (K_TAIT | K_MORRIS)) && rsq < T[I_FLUID].cutsq[ij]) {
        const PairTab &P = T[I_FLUID];
        double h = P.h[ij];
        double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;              // Lucy (dW/dr)/r  (:135-151)
        double fpair, fvisc;
        if (KINDS & K_TAIT) {
          fvisc = 0.0;
          if (dvdr < 0.0) {                                           // Monaghan artificial viscosity (:163-169)
            double mu = h * dvdr / (rsq + 0.01 * h * h);
            fvisc = -P.visc[ij] * (P.cs[ti] + P.cs[tj]) * mu / (rhoi + rhoj);
          }
          fpair = -mi * mj * (qi.x + qj.x + fvisc) * wfd;
          fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
          ade += -0.5 * fpair * dvdr;
        } else {                                                      // Morris viscosity (morris :165-176)
          fvisc = 2.0 * P.visc[ij] / (rhoi * rhoj);
          fvisc *= mi * mj * wfd;
          fpair = -mi * mj * (qi.x + qj.x) * wfd;
          fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
          ade += -0.5 * (fpair * dvdr + fvisc * (dvx * dvx + dvy * dvy + dvz * dvz));
        }
        adrho += mj * dvdr * wfd;
      }
      if ((KINDS & K_HEAT) && rsq < T[I_HEAT].cutsq[ij]) {
        const PairTab &P = T[I_HEAT];
        double h = P.h[ij];
        double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;
        double dE = 2.0 * mi * mj / (mi + mj);
        dE *= (rhoi + rhoj) / (rhoi * rhoj);
        dE *= P.visc[ij] * (qi.y - qj.y) * wfd;                        // D (e_i - e_j) W'/r  (:122-125)
        ade += dE;
      }
      if (MP) {
        double rinv = 1.0 / r;
        if ((KINDS & K_TAITMP) && rsq < T[I_FLUID].cutsq[ij]) {
          const PairTab &P = T[I_FLUID];
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;   // (dW/dr)/r (:137-143)
          double Pi = qi.x, Pj = qj.x;
          if (!P.gamma_uniform) {                                      // p_j uses gamma of the list owner (:148)
            int to = row_owns ? ti : tj;
            Pi = P.B[ti] * (pow(rhoi / P.rho0[ti], P.gamma[to]) - P.rb[ti]);
            Pj = P.B[tj] * (pow(rhoj / P.rho0[tj], P.gamma[to]) - P.rb[tj]);
          }
          double pij = (rhoj * Pi + rhoi * Pj) / (rhoi + rhoj);
          double V2 = qi.y + qj.y;
          double fvisc = V2 * P.visc[ij] * wfd, fpair = -V2 * pij * wfd;
          fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
        }
        if ((KINDS & K_SURF) && rsq < T[I_SURF].cutsq[ij]) {
          const PairTab &P = T[I_SURF];
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij];    // dW/dr (:117-123)
          double ex = dx * rinv, ey = dy * rinv, ez = dz * rinv;
          double six, siy, siz = 0.0, sjx, sjy, sjz = 0.0;
          if (DIM3) {                                                  // (:153-169)
            const double o3 = 0.3333333333333333, t3 = 0.6666666666666666;
            double cxx = ci.x * ci.x, cyy = ci.y * ci.y, czz = ci.z * ci.z;
            six = (ex * (o3 * czz + o3 * cyy - t3 * cxx) - ci.x * ez * ci.z - ci.x * ey * ci.y) * qi.w;
            siy = (ey * (o3 * czz - t3 * cyy + o3 * cxx) - ci.y * ez * ci.z - ex * ci.x * ci.y) * qi.w;
            siz = (ez * (-t3 * czz + o3 * cyy + o3 * cxx) - ey * ci.y * ci.z - ex * ci.x * ci.z) * qi.w;
            cxx = cj.x * cj.x; cyy = cj.y * cj.y; czz = cj.z * cj.z;
            sjx = (ex * (o3 * czz + o3 * cyy - t3 * cxx) - cj.x * ez * cj.z - cj.x * ey * cj.y) * qj.w;
            sjy = (ey * (o3 * czz - t3 * cyy + o3 * cxx) - cj.y * ez * cj.z - ex * cj.x * cj.y) * qj.w;
            sjz = (ez * (-t3 * czz + o3 * cyy + o3 * cxx) - ey * cj.y * cj.z - ex * cj.x * cj.z) * qj.w;
          } else {                                                     // (:140-151); 1/|cg| uses the 2-D norm there
            double ni = sqrt(ci.x * ci.x + ci.y * ci.y), nj = sqrt(cj.x * cj.x + cj.y * cj.y);
            double ii = ni > EPSILON_CG ? 1.0 / ni : 0.0, ij2 = nj > EPSILON_CG ? 1.0 / nj : 0.0;
            double hi2 = (ci.y * ci.y + ci.x * ci.x) / 2, hj2 = (cj.y * cj.y + cj.x * cj.x) / 2;
            six = (ex * (hi2 - ci.x * ci.x) - ci.x * ey * ci.y) * ii;
            siy = (ey * (hi2 - ci.y * ci.y) - ex * ci.x * ci.y) * ii;
            sjx = (ex * (hj2 - cj.x * cj.x) - cj.x * ey * cj.y) * ij2;
            sjy = (ey * (hj2 - cj.y * cj.y) - ex * cj.x * cj.y) * ij2;
          }
          fx += (six * qi.y + sjx * qj.y) * wfd;
          fy += (siy * qi.y + sjy * qj.y) * wfd;
          if (DIM3) fz += (siz * qi.y + sjz * qj.y) * wfd;
        }
        if ((KINDS & (K_HEATMP | K_HEATPC)) && rsq < T[I_HEAT].cutsq[ij]) {
          const PairTab &P = T[I_HEAT];
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;
          double Ti = qi.z, Tj = qj.z;
          if (KINDS & K_HEATPC) {                                      // (:124-129), in half-list orientation
            int ff = P.fixflag[ij]; double tc = P.tc[ij];
            double Ta = row_owns ? Ti : Tj, Tb = row_owns ? Tj : Ti;
            int ta = row_owns ? ti : tj, tb = row_owns ? tj : ti;
            if (ff == ta && Ta < Tb) Ta = tc;
            if (ff == tb && Tb < Ta) Tb = tc;
            Ti = row_owns ? Ta : Tb; Tj = row_owns ? Tb : Ta;
          }
          double dE = 2.0 * P.visc[ij] * (Ti - Tj) * wfd / (rhoi * rhoj);
          ade += dE * mj;
        }
      }
    };
    for_each_hit<true>(A.nbr + (size_t)i * A.stride, A.numneigh[i], lane, queue[wib], test, proc);
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (KINDS & (K_TAIT | K_MORRIS)) adrho = warp_sum(adrho);
    if (KINDS & ~(K_TAITMP | K_SURF)) ade = warp_sum(ade);
    if (lane == 0) {
      double4 f = A.fd[i];
      f.x += fx; f.y += fy; f.z += fz; f.w += adrho;
      A.fd[i] = f;
      if (KINDS & ~(K_TAITMP | K_SURF)) A.de[i] += ade;
    }
  }
}
