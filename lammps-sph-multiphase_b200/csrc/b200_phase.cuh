// b200_phase.cuh -- fix phase_change on the device.
// Restates FixPhaseChange::pre_exchange, /root/reference/src/USER-SPH/fix_phase_change.cpp:167-352
// (+ isfromphasearound :540-563, create_newpos :473-513, create_newpos_simple :467-471,
//  insert_one_atom :425-465, RanPark::uniform src/random_park.cpp:40-47).
//
// The reference walks the owned atoms serially in local order and consumes one RNG draw per
// candidate, plus 2-3 draws per placement attempt of an accepted candidate.  Everything that does
// not depend on the RNG stream position is computed in parallel first (k_pc_candidates: the
// candidate test, the probability threshold and the from-phase-neighbour test); the stream itself
// is then walked by ONE warp in the reference's order (k_pc_walk), which also performs the
// insertions one after the other so that the mass debits land in a fixed order.
//
// Deliberate waiver: the reference writes each new atom over the first live ghost
// slots (create_atom at index nlocal) while later candidates of the same call still read those
// ghosts.  The engine keeps new atoms in a staging area instead; results differ only if one of the
// first `nins` ghosts of the reference's border order is a from-phase neighbour of a later candidate.
#pragma once
#include "b200_common.cuh"
#include "b200_pair.cuh"
#include "b200_tile.cuh"

#define CG_SMALL 1.0e-20
#define PC_MAXNEW 262144

struct PcParams {
  b200_phase_change_desc d;
  int dim, nlocal, nall, stride, norig;
  double dt;
  double sublo[3], subhi[3], boxhi[3];
};
struct PcArrays {
  double4 *xt, *vr, *vm, *cgm;
  double *e, *cv;
  const int *orig;
  const unsigned *nbr, *far;
  const int *numneigh, *numfar;
  // tile path (b200_tile.cuh): the rows hold 16-bit slot entries of the row's tile
  int tiled, ngrp; const TileDesc *tiles; const int *rowtile, *gorder;
};
// the neighbor row of owned atom i, whichever path built it: positions 0..n-1, some of which may be empty
struct PcRow {
  const PcArrays &a; int i, nlocal, stride;
  int nin, nout, nfar, n;
  const unsigned *p, *pf; const unsigned short *tn, *tf; const TileDesc *D;
  __device__ __forceinline__ PcRow(const PcArrays &a_, const PcParams &P, int i_) : a(a_), i(i_), nlocal(P.nlocal), stride(P.stride)
  {
    if (a.tiled) {
      const size_t rbase = (size_t)(i >> 5) * a.ngrp * 32 + (i & 31);
      tn = (const unsigned short *)((const uint4 *)a.nbr + rbase); tf = (const unsigned short *)((const uint4 *)a.far + rbase);
      nin = ((a.numneigh[i] + 7) >> 3) * 8; nout = (((a.numfar[i] >> 16) + 7) >> 3) * 8; nfar = (((a.numfar[i] & 0xffff) + 7) >> 3) * 8;   // nout: the mid zone, stored from the back of the far row
      D = a.tiles + a.rowtile[i];
    } else {
      int c = a.numneigh[i]; nin = c & 0xffff; nout = c >> 16; nfar = a.numfar[i];
      p = row_base(a.nbr, i, stride); pf = row_base(a.far, i, stride);
    }
    n = nin + nout + nfar;
  }
  // entry k: false if the position is empty; j = device index, tj = type
  __device__ __forceinline__ bool get(int k, int &j, int &tj) const
  {
    if (a.tiled) {
      unsigned ent;
      if (k < nin) ent = tn[(size_t)(k >> 3) * 256 + (k & 7)];
      else if (k < nin + nout) { k -= nin; ent = tf[(size_t)(a.ngrp - 1 - (k >> 3)) * 256 + (k & 7)]; }
      else { k -= nin + nout; ent = tf[(size_t)(k >> 3) * 256 + (k & 7)]; }
      if (!ent) return false;
      tj = ent >> TILE_SLOT_BITS; j = tile_slot_src(*D, ent & TMP_SLOT_MASK, nlocal, a.gorder);
      return true;
    }
    unsigned ent = k < nin ? p[(size_t)k * 32] : (k < nin + nout ? p[(size_t)(stride - 1 - (k - nin)) * 32] : pf[(size_t)(k - nin - nout) * 32]);
    tj = (ent >> NBR_TYPE_SHIFT) & 7; j = ent & NBR_INDEX_MASK;
    return true;
  }
};
struct PcNew { double x[3], v[3], vest[3], rho, cv, e; int parent; int pad; };

__device__ __forceinline__ double pc_kernel_quintic(int dim, double r)
{ // sph_kernel_quintic{2d,3d}(r), sph_kernel_quintic.cpp:17-43
  const double norm = dim == 3 ? 0.0716197243913529 : 0.04195297663091802;
  return norm * quintic_w(3.0 * r);
}
__device__ __forceinline__ double ranpark(int *seed)
{
  int k = *seed / 127773;
  *seed = 16807 * (*seed - k * 127773) - 2836 * k;
  if (*seed < 0) *seed += 2147483647;
  return (1.0 / 2147483647) * *seed;
}

// RanPark is the Park-Miller generator seed' = 16807 seed mod (2^31 - 1) (Schrage's split in ranpark() gives the same integers), so
// k steps at once are one modular multiplication by 16807^k: the walk below evaluates the draws of 32 candidates in parallel.
__device__ __forceinline__ int ranpark_skip(unsigned apow, int seed) { return (int)(((unsigned long long)apow * (unsigned)seed) % 2147483647ull); }

// one thread per owned atom (device order): writes, indexed by LAMMPS local index,
//   flag = 0 not a candidate | 1 candidate | 3 candidate that passes every non-random condition
//   thr  = probability threshold of the draw, dev = device index
__global__ void k_pc_candidates(PcParams P, PcArrays a, unsigned char *flag, double *thr, int *dev, double *dmass)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.nall) dmass[i] = 0.0;
  if (i >= P.nlocal) return;
  const b200_phase_change_desc &d = P.d;
  int o = a.orig[i];
  double4 x = a.xt[i];
  int type = tw_type(__double_as_longlong(x.w));
  double e = a.e[i], cv = a.cv[i], Ti = e / cv;
  dev[o] = i;
  if ((Ti < d.Tc) || (type != d.to_type)) { flag[o] = 0; return; }
  bool ok; double t;
  if (d.energy_chance_flag) { t = (e - cv * d.Tc) / d.Hwv * P.dt * d.phase_change_rate; ok = true; }
  else { t = d.change_chance; ok = Ti > d.Tt; }
  if (ok) {   // isfromphasearound(i): any from_type neighbour of the LAST build within the fix cutoff (rsq <= cutoff^2)
    bool around = false;
    double cutoff2 = d.cutoff * d.cutoff;
    PcRow R(a, P, i);
    for (int k = 0; k < R.n && !around; k++) {
      int j, tj;
      if (!R.get(k, j, tj) || tj != d.from_type) continue;
      double4 xj = a.xt[j];
      double rsq = rsq_nofma(x.x - xj.x, x.y - xj.y, x.z - xj.z);
      if (rsq <= cutoff2) around = true;
    }
    ok = around;
  }
  thr[o] = t;
  flag[o] = ok ? 3 : 1;
}

// candidates in ascending local index: flag != 0 -> position (exclusive scan done by the host) -> compact list
__global__ void k_pc_mark(int norig, const unsigned char *flag, int *pos)
{ int o = blockIdx.x * blockDim.x + threadIdx.x; if (o < norig) pos[o] = flag[o] != 0; }
__global__ void k_pc_compact(int norig, const unsigned char *flag, const int *pos, int *list)
{ int o = blockIdx.x * blockDim.x + threadIdx.x; if (o < norig && flag[o]) list[pos[o]] = o; }

// ONE warp walks the RNG stream in LAMMPS local order and performs the insertions.
// state[0] = RanPark seed (persistent), state[1] = number of insertions of this call (out)
__global__ void __launch_bounds__(32) k_pc_walk(PcParams P, PcArrays a, const unsigned char *flag, const double *thr, const int *dev,
                                                double *dmass, PcNew *newatoms, int *state, const int *clist, int ncand)
{
  const b200_phase_change_desc &d = P.d;
  const int lane = threadIdx.x;
  int seed = state[0], nins = 0;
  // lane r holds 16807^(r+1) mod (2^31 - 1): the multiplier that advances the stream by r + 1 draws
  unsigned apow = 16807u;
  for (int k = 0; k < lane; k++) apow = (unsigned)(((unsigned long long)apow * 16807u) % 2147483647ull);
  for (int base = 0; base < ncand; base += 32) {
    int oc_l = base + lane < ncand ? clist[base + lane] : 0;
    unsigned char fl = base + lane < ncand ? flag[oc_l] : 0;
    double thr_l = base + lane < ncand ? thr[oc_l] : 0.0;
    unsigned cand = __ballot_sync(FULLMASK, fl != 0);
    while (cand) {
      // The draw happens for every candidate, before the other tests (:210,212), one draw each as long as nobody passes.  All remaining
      // candidates of this batch take theirs at once: candidate number r (0-based among the remaining ones) sees the stream r + 1 steps on.
      // Only a candidate that passes consumes more (its position draws); the batch is then resumed behind it with the stream where that left it.
      const int r = __popc(cand & ((1u << lane) - 1u));
      const unsigned mult = __shfl_sync(FULLMASK, apow, r);
      const double u_l = (1.0 / 2147483647) * ranpark_skip(mult, seed);
      const unsigned pass = __ballot_sync(FULLMASK, ((cand >> lane) & 1u) && fl == 3 && u_l < thr_l);
      if (!pass) {                                   // nobody: the stream moves on by the number of candidates
        seed = ranpark_skip(__shfl_sync(FULLMASK, apow, __popc(cand) - 1), seed);
        break;
      }
      const int b = __ffs(pass) - 1;                 // the first one that passes, in LAMMPS local order
      const unsigned upto = cand & ((2u << b) - 1u);
      seed = ranpark_skip(__shfl_sync(FULLMASK, apow, __popc(upto) - 1), seed);
      cand &= ~upto;
      int oc = __shfl_sync(FULLMASK, oc_l, b);
      int i = dev[oc];
      double4 xi = a.xt[i], ci = a.cgm[i];
      double coord[3]; bool ok = false;
      for (int phase = 0; phase < 2 && !ok; phase++) {
        double delta = d.dr;
        for (int at = 0; at < d.maxattempt && !ok; at++) {
          if (phase == 0) {                          // create_newpos (:473-513)
            double eij[3];
            if (P.dim == 3) {
              double b1[3] = {-ci.y, ci.x, 0.0};
              double b1abs = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
              if (b1abs > CG_SMALL) { b1[0] /= b1abs; b1[1] /= b1abs; b1[2] /= b1abs; }
              double den = ci.y * ci.y + ci.x * ci.x;
              double b2[3] = {-ci.x * ci.y * ci.z / den, -ci.z * (ci.y * ci.y) / den, ci.y};
              double b2abs = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
              if (b1abs > CG_SMALL) { b2[0] /= b2abs; b2[1] /= b2abs; b2[2] /= b2abs; }
              double atmp = ranpark(&seed) - 0.5;
              double btmp = ranpark(&seed) - 0.5;
              for (int q = 0; q < 3; q++) eij[q] = atmp * b1[q] + btmp * b2[q];
            } else {
              double atmp = ranpark(&seed);
              atmp = atmp > 0.5 ? 1.0 : -1.0;
              eij[0] = -atmp * ci.y; eij[1] = atmp * ci.x; eij[2] = 0.0;
            }
            double eabs = sqrt(eij[0] * eij[0] + eij[1] * eij[1] + eij[2] * eij[2]);
            coord[0] = xi.x + eij[0] * delta / eabs; coord[1] = xi.y + eij[1] * delta / eabs; coord[2] = xi.z + eij[2] * delta / eabs;
          } else {                                   // create_newpos_simple (:467-471)
            coord[0] = xi.x + (ranpark(&seed) - 0.5) * delta;
            coord[1] = xi.y + (ranpark(&seed) - 0.5) * delta;
            coord[2] = xi.z + (ranpark(&seed) - 0.5) * delta;
          }
          // insert_one_atom's sub-box test (:444-453)
          bool in01 = coord[0] >= P.sublo[0] && coord[0] < P.subhi[0] && coord[1] >= P.sublo[1] && coord[1] < P.subhi[1];
          if (in01 && coord[2] >= P.sublo[2] && coord[2] < P.subhi[2]) ok = true;
          else if (P.dim == 3 && coord[2] >= P.boxhi[2] && in01) ok = true;
          else if (P.dim == 2 && coord[1] >= P.boxhi[1] && coord[0] >= P.sublo[0] && coord[0] < P.subhi[0]) ok = true;
          delta = 0.75 * delta;
        }
      }
      if (!ok) continue;
      if (nins >= PC_MAXNEW) { state[2] = 1; continue; }
      // mass / momentum taken from from_type neighbours, weights w = W_quintic(r * cutoff) (:252-300)
      PcRow R(a, P, i);
      const int ntot = R.n;
      double wtot = 0.0;
      for (int k = lane; k < ntot; k += 32) {
        int j, tj;
        if (R.get(k, j, tj) && tj == d.from_type && a.vm[j].w > 0.5 * d.to_mass) {
          double4 xj = a.xt[j];
          double rsq = rsq_nofma(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
          wtot += pc_kernel_quintic(P.dim, sqrt(rsq) * d.cutoff);
        }
      }
#pragma unroll
      for (int s = 16; s; s >>= 1) wtot += __shfl_xor_sync(FULLMASK, wtot, s);
      double mom[6] = {0, 0, 0, 0, 0, 0};
      for (int k = lane; k < ntot; k += 32) {
        int j, tj;
        if (!R.get(k, j, tj)) continue;
        double4 vj = a.vm[j];
        if (tj == d.from_type && vj.w > 0.5 * d.to_mass) {
          double4 xj = a.xt[j], vej = a.vr[j];
          double rsq = rsq_nofma(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
          double dm = d.to_mass * pc_kernel_quintic(P.dim, sqrt(rsq) * d.cutoff) / wtot;
          dmass[j] += dm;                             // entries of one row are distinct atoms; insertions are sequential
          mom[0] += vj.x * dm; mom[1] += vj.y * dm; mom[2] += vj.z * dm;
          mom[3] += vej.x * dm; mom[4] += vej.y * dm; mom[5] += vej.z * dm;
        }
      }
#pragma unroll
      for (int q = 0; q < 6; q++)
#pragma unroll
        for (int s = 16; s; s >>= 1) mom[q] += __shfl_xor_sync(FULLMASK, mom[q], s);
      __syncwarp();
      if (lane == 0) {
        PcNew n;
        double ei = a.e[i], eaux = 0.5 * (ei - d.Hwv);  // conserve energy (:314-317)
        for (int q = 0; q < 3; q++) { n.x[q] = coord[q]; n.v[q] = mom[q] / d.to_mass; n.vest[q] = mom[3 + q] / d.to_mass; }
        n.rho = a.vr[i].w; n.cv = a.cv[i]; n.e = eaux; n.parent = i; n.pad = 0;
        a.e[i] = eaux;
        newatoms[nins] = n;
      }
      nins++;
      __syncwarp();
    }
  }
  if (lane == 0) { state[0] = seed; state[1] = nins; }
}

// (comm->reverse_comm_fix has already added the ghosts' dmass to their owners) the mass debit / energy renormalisation (:324-332)
__global__ void k_pc_apply(int nlocal, PcArrays a, const double *dmass)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double dm = dmass[i];
  double4 v = a.vm[i];
  double mold = v.w;
  v.w = mold - dm;
  a.vm[i] = v;
  double4 c = a.cgm[i]; c.w = v.w; a.cgm[i] = c;
  a.e[i] = a.e[i] * mold / v.w;
}

// append the new atoms behind the owned ones (AtomVecMesoMultiPhase::create_atom defaults,
// atom_vec_meso_multiphase.cpp:968-997, then :301-317); tags = maxtag+1.. in creation order (Atom::tag_extend)
struct AppendArrays { double4 *xt, *vr, *vm, *fd, *cgm; double *e, *de, *cv; int *tag, *mask, *orig; };
__global__ void k_pc_append(int nlocal, int nins, const PcNew *newatoms, AppendArrays a, int to_type, double to_mass, int groupbit, int maxtag, int orig0)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nins) return;
  const PcNew &n = newatoms[k];
  int m = nlocal + k;
  a.xt[m] = make_double4(n.x[0], n.x[1], n.x[2], __longlong_as_double((long long)pack_tw(to_type, 0, 0, 0)));
  a.vm[m] = make_double4(n.v[0], n.v[1], n.v[2], to_mass);
  a.vr[m] = make_double4(n.vest[0], n.vest[1], n.vest[2], n.rho);
  a.fd[m] = make_double4(0, 0, 0, 0);
  a.cgm[m] = make_double4(0, 0, 0, to_mass);
  a.e[m] = n.e; a.de[m] = 0.0; a.cv[m] = n.cv;
  a.tag[m] = maxtag + 1 + k; a.mask[m] = 1 | groupbit; a.orig[m] = orig0 + k;
}
