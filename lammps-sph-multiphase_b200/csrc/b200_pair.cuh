// b200_pair.cuh -- the pair-style stages: one thread per neighbor row.
//
// Each kernel restates one (or a fused group of) PairSPH*::compute loops of
// /root/reference/src/USER-SPH (file:line cited at each formula).
//
// Data path (chosen from the round-1 ncu profiles, profiles/r01_*): the first version
// (one warp per row, ballot compaction, shuffle reduction) issued on 26 % of the cycles and
// moved 31 GB through L1 per force launch.  This version
//   * stores the rows interleaved by 32 (entry k of row r at [(r/32)*stride*32 + k*32 + r%32]),
//     so a warp = 32 consecutive (cell-sorted) particles reads its k-th entries as one 128-byte line;
//   * gathers ONE contiguous record per neighbor (64 B single-phase, 128 B multiphase, written by
//     k_records just before the pass) with 256-bit loads (LDG.E.ENL2.256 on sm_100a) instead of
//     3-4 separate 32-byte arrays;
//   * walks rows that k_build split into an inner zone (inside the pair cutoff at build time) and an
//     outer zone (skin shell / exactly on the cutoff): every entry is still tested each step, but hits
//     and misses are clustered, so a warp does not diverge on the heavy pair body (a per-step pruned
//     copy of the list was tried first: its scattered 4-byte stores cost as much as the force pass);
//   * keeps the partial sums of a particle in one thread: no shuffles, no atomics, and a summation
//     order that is a pure function of the row.
// The type of j travels in the list entry (3 bits), so the tests need no second gather.
//
// The reference visits every pair once (half list, newton on) and updates both atoms; here every
// row evaluates its own side of each pair.  All pair formulas are bitwise symmetric under i<->j
// except two quirks, which use the half-list ownership bit frozen in the entry at build time:
//   - sph/taitwater/multiphase evaluates p_j with gamma[itype]   (pair_sph_taitwater_multiphase.cpp:148)
//   - sph/heatconduction/phasechange clamps Ti then Tj sequentially (..._phasechange.cpp:124-129)
// Pairs with a ghost are evaluated exactly where the reference evaluates them (the side whose half
// list holds the pair, neigh_derive.cpp:121-134) because ghost rho / colorgradient can be one step
// stale in the multiphase styles; ghost rows accumulate what the
// reference adds to ghost atoms before reverse_comm.
#pragma once
#include "b200_common.cuh"
#include "b200_neigh.cuh"

#define PAIR_THREADS 128
#define EPSILON_CG 1.0e-12   // pair_sph_surfacetension.cpp EPSILON

__device__ __forceinline__ double4 ld256(const double4 *p)
{
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double4 *p, const double4 &v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ const unsigned *row_base(const unsigned *list, int row, int stride)
{
  return list + (size_t)(row >> 5) * stride * 32 + (row & 31);
}
__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULLMASK, v, o));
  return v;
}
__device__ __forceinline__ void load_tab(PairTab *dst, const PairTab *src)
{
  for (int k = threadIdx.x; k < (int)(sizeof(PairTab) / 4); k += blockDim.x) ((int *)dst)[k] = ((const int *)src)[k];
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

struct PairArgs {
  int nlocal, nall, stride, dim, multiphase, nrec;
  const unsigned *list, *far;  // near rows (inner + outer zone) and far rows
  const int *cnt, *numfar, *scan_far;
  const double4 *xt, *vr, *vm, *cgm;
  const double *e, *cv;
  double4 *rec;                // gather records, nrec double4 per particle
  double4 *vr_out, *cg_out, *fd;
  double *de;
  const PairTab *tab[4];
};

// ---- walking a row: inner zone front to back, then the outer zone (k_build) ----
struct RowIter {
  const unsigned *p, *pf; int n_in, n_out, n_far, m_in, m_near, m_tot, last;
  __device__ __forceinline__ RowIter(const unsigned *list, const int *cnt, const unsigned *far, const int *numfar, const int *scan_far,
                                     int row, int stride, bool valid)
  {
    int c = valid ? cnt[row] : 0;
    n_in = c & 0xffff; n_out = c >> 16;
    n_far = (valid && *scan_far) ? numfar[row] : 0;      // far rows only once a particle has moved margin/2 since the build
    m_in = warp_max(n_in); m_near = m_in + warp_max(n_out); m_tot = m_near + warp_max(n_far);
    p = row_base(list, row, stride); pf = row_base(far, row, stride); last = stride - 1;
  }
  // entry of step k (0 <= k < m_tot), or false if this lane has none at that step
  __device__ __forceinline__ bool get(int k, unsigned &ent) const
  {
    const unsigned *q; bool has;
    if (k < m_in) { q = p + (size_t)k * 32; has = k < n_in; }
    else if (k < m_near) { q = p + (size_t)(last - (k - m_in)) * 32; has = k - m_in < n_out; }
    else { q = pf + (size_t)(k - m_near) * 32; has = k - m_near < n_far; }
    if (has) ent = __ldg(q);
    return has;
  }
};

// ------------------------------------------------------------------ density --
// MP=false: PairSPHRhoSum::compute            pair_sph_rhosum.cpp:112-197   (quadric kernel, per-type mass)
// MP=true : PairSPHRhoSumMultiphase::compute  pair_sph_rhosum_multiphase.cpp:113-168 (quintic, number density * own mass)
template <bool MP>
__global__ void __launch_bounds__(PAIR_THREADS) k_rhosum(PairArgs A)
{
  __shared__ PairTab T;
  load_tab(&T, A.tab[0]);
  __syncthreads();
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = row < A.nlocal;
  double4 pi = valid ? A.xt[row] : make_double4(0, 0, 0, 0);
  int ti = tw_type(__double_as_longlong(pi.w));
  if (valid && T.iskip[ti]) valid = false;           // atoms of skipped types keep their integrated rho
  RowIter it(A.list, A.cnt, A.far, A.numfar, A.scan_far, row, A.stride, valid);
  if (!it.m_tot && !valid) return;
  double acc = 0.0;
#pragma unroll 4
  for (int k = 0; k < it.m_tot; k++) {
    unsigned ent;
    if (it.get(k, ent)) {
      double4 pj = ld256(A.xt + (ent & NBR_INDEX_MASK));
      int tj = (ent >> NBR_TYPE_SHIFT) & 7, ij = ti * MAXT1 + tj;
      double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      if (rsq < T.cutsq[ij]) {
        if (!MP) {
          double wf = 1.0 - rsq * T.c1[ij];            // 1 - r^2/h^2
          wf = wf * wf; wf = wf * wf;
          acc += T.mass[tj] * (T.c0[ij] * wf);         // C_d (1-r^2/h^2)^4 / h^d
        } else {
          acc += T.c0[ij] * quintic_w(3.0 * (sqrt(rsq) * T.c1[ij]));
        }
      }
    }
  }
  if (valid) {
    double rho;
    if (!MP) rho = T.mass[ti] * T.self0[ti] + acc;
    else rho = (T.self0[ti] + acc) * A.vm[row].w;      // rho[i] *= imass (:170)
    A.vr_out[row].w = rho;
  }
}

// ------------------------------------------------------------------ records --
// Per-pass gather records for all nall particles (ghosts keep their possibly stale rho / cg):
//  mode 0 (single-phase force): r0 = x,y,z,rho   r1 = vest, Tait term B((rho/rho0)^7-1)/rho^2 (pair_sph_taitwater.cpp:118-120)
//                               r2 = e,-,-,-     (only when nrec == 3; heat alone: r1 = e,-,-,-)
//  mode 1 (multiphase force)  : r0 = x,y,z,rho   r1 = vest, pressure (pair_sph_taitwater_multiphase.cpp:289-292)
//                               r2 = cg, 1/|cg| or 0    r3 = V^2 = (m/rho)^2, T = e/cv, m, -
//  mode 2 (colorgradient)     : r0 = x,y,z, V^2 = 1/sigma^2
__global__ void k_records(int nall, int mode, int nrec, int heat_only, const PairTab *fluid, const double4 *xt, const double4 *vr,
                          const double4 *cgm, const double *e, const double *cv, double4 *rec)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nall) return;
  double4 x = xt[i], v = vr[i];
  int t = tw_type(__double_as_longlong(x.w));
  double rho = v.w;
  double4 *r = rec + (size_t)i * nrec;
  if (mode == 2) {
    double V = cgm[i].w / rho;
    st256(r, make_double4(x.x, x.y, x.z, V * V));
    return;
  }
  st256(r, make_double4(x.x, x.y, x.z, rho));
  if (mode == 0) {
    double pf = 0.0;
    if (fluid && fluid->style == B200_PAIR_IDEALGAS) pf = 0.4 * e[i] / fluid->mass[t] / rho;      // p / rho^2, pair_sph_idealgas.cpp:94
    else if (fluid) {
      double tmp = rho / fluid->rho0[t], fi = tmp * tmp * tmp;
      pf = fluid->B[t] * (fi * fi * tmp - 1.0) / (rho * rho);
    }
    if (heat_only) st256(r + 1, make_double4(e[i], 0.0, 0.0, 0.0));
    else st256(r + 1, make_double4(v.x, v.y, v.z, pf));
    if (nrec == 3) st256(r + 2, make_double4(e[i], 0.0, 0.0, 0.0));
  } else {
    double4 c = cgm[i];
    double P = fluid ? fluid->B[t] * (pow(rho / fluid->rho0[t], fluid->gamma[t]) - fluid->rb[t]) : 0.0;
    double a = sqrt(c.x * c.x + c.y * c.y + c.z * c.z);
    double V = c.w / rho;
    st256(r + 1, make_double4(v.x, v.y, v.z, P));
    st256(r + 2, make_double4(c.x, c.y, c.z, a > EPSILON_CG ? 1.0 / a : 0.0));
    st256(r + 3, make_double4(V * V, e[i] / cv[i], c.w, 0.0));
  }
}

// PairSPHColorGradient::compute, pair_sph_colorgradient.cpp:119-184 (full list, owned rows).
// dphi = -W'(r) alpha sigma_i / sigma_j^2, sigma = rho/m  ->  1/sigma_j^2 = V_j^2 (record .w)
__global__ void __launch_bounds__(PAIR_THREADS) k_colorgradient(PairArgs A)
{
  __shared__ PairTab T;
  load_tab(&T, A.tab[0]);
  __syncthreads();
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = row < A.nlocal;
  double4 pi = valid ? A.xt[row] : make_double4(0, 0, 0, 0);
  int ti = tw_type(__double_as_longlong(pi.w));
  if (valid && T.iskip[ti]) valid = false;
  RowIter it(A.list, A.cnt, A.far, A.numfar, A.scan_far, row, A.stride, valid);
  if (!it.m_tot && !valid) return;
  double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll 2
  for (int k = 0; k < it.m_tot; k++) {
    unsigned ent;
    if (it.get(k, ent)) {
      double4 pj = ld256(A.rec + (ent & NBR_INDEX_MASK));
      int ij = ti * MAXT1 + ((ent >> NBR_TYPE_SHIFT) & 7);
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      double rsq = rsq_nofma(dx, dy, dz);
      if (rsq < T.cutsq[ij]) {
        double rinv = rsqrt(rsq), r = rsq * rinv;
        double wfd = quintic_dw(3.0 * (r * T.c1[ij])) * T.c0[ij];       // dW/dr
        double s = -wfd * T.visc[ij] * pj.w * rinv;                     // -W' alpha / sigma_j^2 / r   (sigma_i applied at the end)
        ax += s * dx; ay += s * dy; az += s * dz;
      }
    }
  }
  if (valid) {
    double sigmai = A.vr[row].w / A.vm[row].w;
    double4 c = A.cg_out[row];
    c.x = ax * sigmai; c.y = ay * sigmai; c.z = (A.dim == 3) ? az * sigmai : 0.0;
    A.cg_out[row] = c;
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
// measured (profiles/r01_force_variants.txt): capping the single-phase kernels at 64 registers (8 blocks of 128 threads per
// SM) gains ~4 %; the same cap on the multiphase kernels spills and loses 6 %; a depth-1 software prefetch of the next
// gather raised the register count and lost 9 %.  The kernel sits between the L1 wavefront limit and gather latency.
#define FORCE_MIN_BLOCKS(K) (((K) & (K_TAITMP | K_SURF | K_HEATMP | K_HEATPC)) ? 1 : 8)
template <int KINDS, bool DIM3>
__global__ void __launch_bounds__(PAIR_THREADS, FORCE_MIN_BLOCKS(KINDS)) k_force(PairArgs A)
{
  constexpr bool HAS_FLUID = (KINDS & (K_TAIT | K_MORRIS | K_TAITMP | K_IDEAL)) != 0;
  constexpr bool HAS_SURF = (KINDS & K_SURF) != 0;
  constexpr bool HAS_HEAT = (KINDS & (K_HEAT | K_HEATMP | K_HEATPC)) != 0;
  constexpr int NK = (HAS_FLUID ? 1 : 0) + (HAS_SURF ? 1 : 0) + (HAS_HEAT ? 1 : 0);
  constexpr bool MP = (KINDS & (K_TAITMP | K_SURF | K_HEATMP | K_HEATPC)) != 0;
  constexpr int NREC = MP ? 4 : ((KINDS & K_HEAT) && HAS_FLUID ? 3 : 2);
  constexpr int I_FLUID = 0;                        // canonical table order: fluid, surf, heat
  constexpr int I_SURF = HAS_FLUID ? 1 : 0;
  constexpr int I_HEAT = I_SURF + (HAS_SURF ? 1 : 0);
  constexpr bool WRITES_DE = (KINDS & ~(K_TAITMP | K_SURF)) != 0;
  __shared__ PairTab T[NK];
  for (int t = 0; t < NK; t++) load_tab(&T[t], A.tab[t]);
  __syncthreads();

  int row = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = row < A.nall;
  RowIter it(A.list, A.cnt, A.far, A.numfar, A.scan_far, row, A.stride, valid);
  if (!it.m_tot) return;                             // fd / de of these rows stay as force_clear left them
  const bool ghostrow = row >= A.nlocal;
  const double4 *ri = A.rec + (size_t)(valid ? row : 0) * NREC;
  double4 p0 = ri[0], p1 = ri[1], p2 = make_double4(0, 0, 0, 0), p3 = p2;
  if (NREC >= 3) p2 = ri[2];
  if (NREC >= 4) p3 = ri[3];
  int ti = valid ? tw_type(__double_as_longlong(A.xt[row].w)) : 1;
  const double rhoi = p0.w;
  const double mi = MP ? p3.z : T[0].mass[ti];
  double fx = 0, fy = 0, fz = 0, adrho = 0, ade = 0;

#pragma unroll 2
  for (int k = 0; k < it.m_tot; k++) {
    unsigned ent;
    if (!it.get(k, ent)) continue;
    int j = ent & NBR_INDEX_MASK;
    if (!ghostrow && j >= A.nlocal && !(ent & NBR_OWNER_BIT)) continue;     // the other side owns this ghost pair
    const bool row_owns = ghostrow ? false : ((ent & NBR_OWNER_BIT) != 0);   // is the row particle the reference's "i" of this pair?
    const double4 *rj = A.rec + (size_t)j * NREC;
    double4 q0 = ld256(rj);
    int tj = (ent >> NBR_TYPE_SHIFT) & 7, ij = ti * MAXT1 + tj;
    double dx = p0.x - q0.x, dy = p0.y - q0.y, dz = p0.z - q0.z;
    double rsq = rsq_nofma(dx, dy, dz);
    bool any = false;
#pragma unroll
    for (int t = 0; t < NK; t++) any |= rsq < T[t].cutsq[ij];
    if (!any) continue;
    double4 q1 = ld256(rj + 1), q2 = make_double4(0, 0, 0, 0), q3 = q2;
    if (NREC >= 3) q2 = ld256(rj + 2);
    if (NREC >= 4) q3 = ld256(rj + 3);
    const double rhoj = q0.w;
    const double mj = MP ? q3.z : T[0].mass[tj];
    // coincident particles (two atoms created at one position, e.g. where the wall and driver regions of cavity_flow.lmp overlap): r = 0 as
    // the reference's sqrt gives, not 0 * inf; the single-phase formulas never divide by r (the multiphase ones do, in the reference as well)
    const double rinv = rsqrt(rsq), r = rsq > 0.0 ? rsq * rinv : 0.0;

    if (KINDS & (K_TAIT | K_MORRIS | K_IDEAL)) {
      const PairTab &P = T[I_FLUID];
      if (rsq < P.cutsq[ij]) {
        double h = P.h[ij];
        double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;              // Lucy (dW/dr)/r  (:135-151)
        double dvx = p1.x - q1.x, dvy = p1.y - q1.y, dvz = p1.z - q1.z;
        double dvdr = dx * dvx + dy * dvy + dz * dvz;
        double mm = mi * mj;
        if (KINDS & (K_TAIT | K_IDEAL)) {
          double fvisc = 0.0;
          if (dvdr < 0.0) {                                           // Monaghan artificial viscosity (:163-169), one division
            // sph/idealgas: c = sqrt(0.4 e / m) = sqrt((p/rho^2) rho)  (pair_sph_idealgas.cpp:95,141)
            const double cc = (KINDS & K_IDEAL) ? sqrt(p1.w * rhoi) + sqrt(q1.w * rhoj) : P.cs[ti] + P.cs[tj];
            // sph/idealgas never mirrors viscosity[i][j] into [j][i] (pair_sph_idealgas.cpp:244-253): the half-list owner's type comes first
            const int iv = ((KINDS & K_IDEAL) && !row_owns) ? tj * MAXT1 + ti : ij;
            fvisc = -P.visc[iv] * cc * (h * dvdr) / ((rsq + 0.01 * h * h) * (rhoi + rhoj));
          }
          double fpair = -mm * (p1.w + q1.w + fvisc) * wfd;
          fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
          ade += -0.5 * fpair * dvdr;
        } else {                                                      // Morris viscosity (morris :165-176)
          double fvisc = 2.0 * P.visc[ij] / (rhoi * rhoj) * mm * wfd;
          double fpair = -mm * (p1.w + q1.w) * wfd;
          fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
          ade += -0.5 * (fpair * dvdr + fvisc * (dvx * dvx + dvy * dvy + dvz * dvz));
        }
        adrho += mj * dvdr * wfd;
      }
    }
    if (KINDS & K_HEAT) {
      const PairTab &P = T[I_HEAT];
      if (rsq < P.cutsq[ij]) {
        double h = P.h[ij];
        double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;
        double ei = HAS_FLUID ? p2.x : p1.x, ej = HAS_FLUID ? q2.x : q1.x;
        // 2 mi mj/(mi+mj) (rho_i+rho_j)/(rho_i rho_j) D (e_i - e_j) W'/r  (:122-125), one division
        ade += 2.0 * mi * mj * (rhoi + rhoj) * P.visc[ij] * (ei - ej) * wfd / ((mi + mj) * (rhoi * rhoj));
      }
    }
    if (MP) {
      if (KINDS & K_TAITMP) {
        const PairTab &P = T[I_FLUID];
        if (rsq < P.cutsq[ij]) {
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;   // (dW/dr)/r (:137-143)
          double Pi = p1.w, Pj = q1.w;
          if (!P.gamma_uniform) {                                      // p_j uses gamma of the list owner (:148)
            int to = row_owns ? ti : tj;
            Pi = P.B[ti] * (pow(rhoi / P.rho0[ti], P.gamma[to]) - P.rb[ti]);
            Pj = P.B[tj] * (pow(rhoj / P.rho0[tj], P.gamma[to]) - P.rb[tj]);
          }
          double pij = (rhoj * Pi + rhoi * Pj) / (rhoi + rhoj);
          double V2 = p3.x + q3.x;
          double fvisc = V2 * P.visc[ij] * wfd, fpair = -V2 * pij * wfd;
          double dvx = p1.x - q1.x, dvy = p1.y - q1.y, dvz = p1.z - q1.z;
          fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
        }
      }
      if (KINDS & K_SURF) {
        const PairTab &P = T[I_SURF];
        if (rsq < P.cutsq[ij]) {
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij];    // dW/dr (:117-123)
          double ex = dx * rinv, ey = dy * rinv, ez = dz * rinv;
          double six, siy, siz = 0.0, sjx, sjy, sjz = 0.0;
          if (DIM3) {                                                  // (:153-169)
            const double o3 = 0.3333333333333333, t3 = 0.6666666666666666;
            double cxx = p2.x * p2.x, cyy = p2.y * p2.y, czz = p2.z * p2.z;
            six = (ex * (o3 * czz + o3 * cyy - t3 * cxx) - p2.x * ez * p2.z - p2.x * ey * p2.y) * p2.w;
            siy = (ey * (o3 * czz - t3 * cyy + o3 * cxx) - p2.y * ez * p2.z - ex * p2.x * p2.y) * p2.w;
            siz = (ez * (-t3 * czz + o3 * cyy + o3 * cxx) - ey * p2.y * p2.z - ex * p2.x * p2.z) * p2.w;
            cxx = q2.x * q2.x; cyy = q2.y * q2.y; czz = q2.z * q2.z;
            sjx = (ex * (o3 * czz + o3 * cyy - t3 * cxx) - q2.x * ez * q2.z - q2.x * ey * q2.y) * q2.w;
            sjy = (ey * (o3 * czz - t3 * cyy + o3 * cxx) - q2.y * ez * q2.z - ex * q2.x * q2.y) * q2.w;
            sjz = (ez * (-t3 * czz + o3 * cyy + o3 * cxx) - ey * q2.y * q2.z - ex * q2.x * q2.z) * q2.w;
          } else {                                                     // (:140-151); |cg| is the 2-D norm there
            double ni = sqrt(p2.x * p2.x + p2.y * p2.y), nj = sqrt(q2.x * q2.x + q2.y * q2.y);
            double ii = ni > EPSILON_CG ? 1.0 / ni : 0.0, jj = nj > EPSILON_CG ? 1.0 / nj : 0.0;
            double hi2 = (p2.y * p2.y + p2.x * p2.x) / 2, hj2 = (q2.y * q2.y + q2.x * q2.x) / 2;
            six = (ex * (hi2 - p2.x * p2.x) - p2.x * ey * p2.y) * ii;
            siy = (ey * (hi2 - p2.y * p2.y) - ex * p2.x * p2.y) * ii;
            sjx = (ex * (hj2 - q2.x * q2.x) - q2.x * ey * q2.y) * jj;
            sjy = (ey * (hj2 - q2.y * q2.y) - ex * q2.x * q2.y) * jj;
          }
          fx += (six * p3.x + sjx * q3.x) * wfd;
          fy += (siy * p3.x + sjy * q3.x) * wfd;
          if (DIM3) fz += (siz * p3.x + sjz * q3.x) * wfd;
        }
      }
      if (KINDS & (K_HEATMP | K_HEATPC)) {
        const PairTab &P = T[I_HEAT];
        if (rsq < P.cutsq[ij]) {
          double wfd = quintic_dw(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;
          double Ti = p3.y, Tj = q3.y;
          if (KINDS & K_HEATPC) {                                      // (:124-129), in half-list orientation
            int ff = P.fixflag[ij]; double tc = P.tc[ij];
            double Ta = row_owns ? Ti : Tj, Tb = row_owns ? Tj : Ti;
            int ta = row_owns ? ti : tj, tb = row_owns ? tj : ti;
            if (ff == ta && Ta < Tb) Ta = tc;
            if (ff == tb && Tb < Ta) Tb = tc;
            Ti = row_owns ? Ta : Tb; Tj = row_owns ? Tb : Ta;
          }
          ade += 2.0 * P.visc[ij] * (Ti - Tj) * wfd / (rhoi * rhoj) * mj;
        }
      }
    }
  }
  if (valid) {
    double4 f = A.fd[row];
    f.x += fx; f.y += fy; f.z += fz; f.w += adrho;
    A.fd[row] = f;
    if (WRITES_DE) A.de[row] += ade;
  }
}
