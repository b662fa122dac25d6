// b200_comm.cuh -- halo (ghost) exchange, reverse accumulation and atom migration kernels.
//
// Restates CommBrick::borders / forward_comm / reverse_comm / exchange
// (/root/reference/src/comm_brick.cpp:444-864) with the AtomVecMeso[MultiPhase] wire formats
// (src/USER-SPH/atom_vec_meso*.cpp pack_border/pack_comm/pack_reverse/pack_exchange), as
// device pack/unpack kernels around an NCCL send/recv (NVLink 5) -- or a plain local copy when
// the neighbour in that direction is this rank itself (periodic self-images).
// Ghosts are stored behind the owned atoms in swap order, exactly like the reference, so each
// swap's receive range is contiguous.
#pragma once
#include "b200_common.cuh"
#include "b200_neigh.cuh"

#define NB_BORDER 21    // x3 v3 vest3 rho cg3 rmass e cv tag type mask image|order-code root-index
#define NB_REVERSE 5    // f3 drho de
#define NB_EXCHANGE 26  // every per-atom field

struct CommArrays {
  double4 *xt, *vr, *vm, *fd, *cgm;
  double *e, *de, *cv;
  int *tag, *mask, *orig, *gimage;
};

// atoms of [0,n) inside the slab lo <= x[dim] <= hi  (comm_brick.cpp:746-750)
__global__ void k_slab_flag(int n, const double4 *xt, int dim, double lo, double hi, int *flag, int *pos)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xt[i];
  double c = dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
  int f = (c >= lo && c <= hi) ? 1 : 0;
  flag[i] = f; pos[i] = f;
}
// both swaps of a dimension in one pass over the atoms (they scan the same atoms, comm_brick.cpp:722-725): flag / pos hold swap 0 in
// [0, n) and swap 1 in [n, 2n), so ONE scan over 2n elements gives pos0 = P, count0 = P[n], pos1[i] = P[n + i] - P[n], count1 = P[2n] - P[n]
__global__ void k_slab_flag2(int n, const double4 *xt, int dim, double lo0, double hi0, int on0, double lo1, double hi1, int on1, int *flag, int *pos)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xt[i];
  double c = dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
  int f0 = (on0 && c >= lo0 && c <= hi0) ? 1 : 0, f1 = (on1 && c >= lo1 && c <= hi1) ? 1 : 0;
  flag[i] = f0; pos[i] = f0; flag[n + i] = f1; pos[n + i] = f1;
}
// bias: where the offsets of this swap start inside a shared scan (NULL = 0)
__global__ void k_compact(int n, const int *flag, const int *pos, int *list, const int *bias)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) list[pos[i] - (bias ? *bias : 0)] = i;
}

__device__ __forceinline__ double shifted(double c, double shift) { return shift != 0.0 ? __dadd_rn(c, shift) : c; }

// gimage of an atom: bits 0-7 the periodic image code (13 = none), bits 8-16 its ORDER CODE, three octal digits d3 d2 d1 = (swap that
// created the ghost) + 1, the same for its source, and for the source's source (0 = an owned atom).  The reference appends ghosts
// swap by swap, and inside a swap in the sender's local-index order (comm_brick.cpp:722-760), whereas the engine scans its
// cell-sorted device order; (order code, orig of the root owned atom) is a sort key that reproduces the reference's ghost order
// wherever the local index of a ghost matters (pair sph/lj: the order of a bin's atoms in the neighbor list, b200_lj.cuh).
__device__ __forceinline__ int ghost_code(int gimage_src, int imgstep, int swapcode)
{
  const int img = (gimage_src & 0xff) + imgstep, d = (gimage_src >> 8) & 0x1ff;
  return img | (((swapcode << 6) | (d >> 3)) << 8);
}

// pack_border_vel: x + pbc shift, ... (atom_vec_meso_multiphase.cpp:555-721)
__global__ void k_pack_border(int n, const int *list, CommArrays a, int dim, double shift, int imgstep, int swapcode, double *buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int j = list[k];
  double4 x = a.xt[j], vr = a.vr[j], v = a.vm[j], c = a.cgm[j];
  double *b = buf + (size_t)k * NB_BORDER;
  b[0] = dim == 0 ? shifted(x.x, shift) : x.x; b[1] = dim == 1 ? shifted(x.y, shift) : x.y; b[2] = dim == 2 ? shifted(x.z, shift) : x.z;
  b[3] = v.x; b[4] = v.y; b[5] = v.z; b[6] = vr.x; b[7] = vr.y; b[8] = vr.z; b[9] = vr.w;
  b[10] = c.x; b[11] = c.y; b[12] = c.z; b[13] = v.w; b[14] = a.e[j]; b[15] = a.cv[j];
  b[16] = (double)a.tag[j]; b[17] = (double)tw_type(__double_as_longlong(x.w)); b[18] = (double)a.mask[j];
  b[19] = (double)ghost_code(a.gimage[j], imgstep, swapcode); b[20] = (double)a.orig[j];
}
__global__ void k_unpack_border(Geom g, int n, int first, CommArrays a, const double *buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int i = first + k;
  const double *b = buf + (size_t)k * NB_BORDER;
  int type = (int)b[17];
  a.xt[i] = make_double4(b[0], b[1], b[2], __longlong_as_double((long long)make_tw(g, type, b[0], b[1], b[2])));
  a.vm[i] = make_double4(b[3], b[4], b[5], b[13]);
  a.vr[i] = make_double4(b[6], b[7], b[8], b[9]);
  a.cgm[i] = make_double4(b[10], b[11], b[12], b[13]);
  a.e[i] = b[14]; a.cv[i] = b[15];
  a.tag[i] = (int)b[16]; a.mask[i] = (int)b[18]; a.gimage[i] = (int)b[19]; a.orig[i] = (int)b[20];      // a ghost's orig = local index of its root owned atom on the sending rank
}
// borders without a host round trip per swap: one thread per atom of [0, nlast); a flagged atom k = pos[i] < cap is packed into
// record k and entered in the send list; buf[0] (the message header) carries the true count pos[nlast], so the receiver -- who
// posted a receive of the same cap, derived from the count of the previous build on both sides -- learns how many records are
// valid, and both sides learn about an overflow (count > cap) from the same number.
__global__ void k_pack_border_compact(int nlast, const int *flag, const int *pos, const int *bias, int cap, int *list, CommArrays a, int dim, double shift,
                                      int imgstep, int swapcode, double *buf)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = bias ? *bias : 0;
  if (i == 0) buf[0] = (double)(pos[nlast] - b0);
  if (i >= nlast || !flag[i]) return;
  int k = pos[i] - b0;
  if (k >= cap) return;
  list[k] = i;
  double4 x = a.xt[i], vr = a.vr[i], v = a.vm[i], c = a.cgm[i];
  double *b = buf + 1 + (size_t)k * NB_BORDER;
  b[0] = dim == 0 ? shifted(x.x, shift) : x.x; b[1] = dim == 1 ? shifted(x.y, shift) : x.y; b[2] = dim == 2 ? shifted(x.z, shift) : x.z;
  b[3] = v.x; b[4] = v.y; b[5] = v.z; b[6] = vr.x; b[7] = vr.y; b[8] = vr.z; b[9] = vr.w;
  b[10] = c.x; b[11] = c.y; b[12] = c.z; b[13] = v.w; b[14] = a.e[i]; b[15] = a.cv[i];
  b[16] = (double)a.tag[i]; b[17] = (double)tw_type(__double_as_longlong(x.w)); b[18] = (double)a.mask[i];
  b[19] = (double)ghost_code(a.gimage[i], imgstep, swapcode); b[20] = (double)a.orig[i];
}
// pack_comm[_vel] (atom_vec_meso.cpp:139-245, atom_vec_meso_multiphase.cpp:319-465): cv, type, tag, mask are NOT resent.
// Message layout (doubles per ghost): x3 vest3 rho e | v3 if comm_modify vel yes | cg3 rmass if atom_style meso/multiphase
// -> 8 for the plain single-phase deck (the reference's own pack_comm size), 15 for the multiphase decks.
__host__ __device__ __forceinline__ int fwd_width(int multiphase, int ghost_velocity) { return 8 + (ghost_velocity ? 3 : 0) + (multiphase ? 4 : 0); }
__global__ void k_pack_forward(int n, const int *list, CommArrays a, int dim, double shift, double *buf, int multiphase, int ghost_velocity)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int j = list[k];
  double4 x = a.xt[j], vr = a.vr[j];
  double *b = buf + (size_t)k * fwd_width(multiphase, ghost_velocity);
  b[0] = dim == 0 ? shifted(x.x, shift) : x.x; b[1] = dim == 1 ? shifted(x.y, shift) : x.y; b[2] = dim == 2 ? shifted(x.z, shift) : x.z;
  b[3] = vr.x; b[4] = vr.y; b[5] = vr.z; b[6] = vr.w; b[7] = a.e[j];
  int o = 8;
  if (ghost_velocity | multiphase) {
    double4 v = a.vm[j];
    if (ghost_velocity) { b[o] = v.x; b[o + 1] = v.y; b[o + 2] = v.z; o += 3; }
    if (multiphase) { double4 c = a.cgm[j]; b[o] = c.x; b[o + 1] = c.y; b[o + 2] = c.z; b[o + 3] = v.w; }
  }
}
// xhold / dmaxsq (track): the ghosts' displacement since the build enters the same bound as the owned atoms' (k_initial_integrate),
// so the far / mid zone flags need no all-reduce: every candidate of this rank's rows is an owned atom or one of its ghosts
__global__ void k_unpack_forward(int n, int first, CommArrays a, const double *buf, int multiphase, int ghost_velocity,
                                 const double *xhold, unsigned long long *dmaxsq, const int *gcell, unsigned *celld, int nlocal)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int i = first + k;
  const double *b = buf + (size_t)k * fwd_width(multiphase, ghost_velocity);
  double4 x = a.xt[i];
  x.x = b[0]; x.y = b[1]; x.z = b[2];
  a.xt[i] = x;
  if (xhold) {
    double dx = x.x - xhold[3 * i], dy = x.y - xhold[3 * i + 1], dz = x.z - xhold[3 * i + 2];
    const double dsq = dx * dx + dy * dy + dz * dz;
    unsigned long long bits = (unsigned long long)__double_as_longlong(dsq);
    if (bits > *(volatile unsigned long long *)dmaxsq) atomicMax(dmaxsq, bits);
    if (celld) cell_disp_max(celld, gcell[i - nlocal], dsq);
  }
  a.vr[i] = make_double4(b[3], b[4], b[5], b[6]);
  a.e[i] = b[7];
  int o = 8;
  if (ghost_velocity | multiphase) {
    double4 v = a.vm[i];
    if (ghost_velocity) { v.x = b[o]; v.y = b[o + 1]; v.z = b[o + 2]; o += 3; }
    if (multiphase) { v.w = b[o + 3]; a.cgm[i] = make_double4(b[o], b[o + 1], b[o + 2], b[o + 3]); }
    a.vm[i] = v;
  }
}
// pack_reverse / unpack_reverse (atom_vec_meso_multiphase.cpp:520-551): ghosts' f, drho, de added to the atoms that were sent
__global__ void k_pack_reverse(int n, int first, CommArrays a, double *buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double4 f = a.fd[first + k];
  double *b = buf + (size_t)k * NB_REVERSE;
  b[0] = f.x; b[1] = f.y; b[2] = f.z; b[3] = f.w; b[4] = a.de[first + k];
}
__global__ void k_unpack_reverse(int n, const int *list, CommArrays a, const double *buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int j = list[k];                       // entries of one sendlist are distinct -> no atomics needed
  const double *b = buf + (size_t)k * NB_REVERSE;
  double4 f = a.fd[j];
  f.x += b[0]; f.y += b[1]; f.z += b[2]; f.w += b[3];
  a.fd[j] = f;
  a.de[j] += b[4];
}
// one double per atom: forward (sph/rhosum's forward_comm_pair, pair_sph_rhosum.cpp:290-313) / reverse-add (fix phase_change dmass)
__global__ void k_pack_rho(int n, const int *list, const double4 *vr, double *buf)
{ int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) buf[k] = vr[list[k]].w; }
__global__ void k_unpack_rho(int n, int first, double4 *vr, const double *buf)
{ int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) vr[first + k].w = buf[k]; }
__global__ void k_pack_scalar(int n, int first, const double *src, double *buf)
{ int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) buf[k] = src[first + k]; }
__global__ void k_unpack_scalar_add(int n, const int *list, double *dst, const double *buf)
{ int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) dst[list[k]] += buf[k]; }

// ---- migration: CommBrick::exchange (comm_brick.cpp:573-684) ----
// atoms outside [lo,hi) in `dim` leave (flag) and are packed; slot is marked dead (mask = 0 convention: alive[] array)
__global__ void k_exchange_flag(int n, const double4 *xt, const int *alive, int dim, double lo, double hi, int *flag, int *pos)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xt[i];
  double c = dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
  int f = (alive[i] && (c < lo || c >= hi)) ? 1 : 0;
  flag[i] = f; pos[i] = f;
}
// ---- CommBrick::exchange's own order (comm_brick.cpp:628-650) ---------------------------------------------------------------
// The reference walks its local indices upwards; an atom that left is packed and the LAST atom is copied into its place
// (avec->copy(nlocal-1,i,1); nlocal--), the same index being examined again.  That decides (a) the order of the atoms in the
// message, hence the local indices they get on the receiving rank, and (b) which of the staying atoms change their index.  The
// local index is the engine's `orig` key (half-list orientation of the two quirks, the RNG walk of fix phase_change, output order),
// so the walk is replayed on the keys: ranks of the keys (dense local indices), the leavers' ranks in ascending order, then one
// thread follows the hole-filling chain -- its length is the number of leavers, not of atoms.
__global__ void k_orig_mark(int n, const int *alive, const int *orig, int *mark) { int s = blockIdx.x * blockDim.x + threadIdx.x; if (s < n && alive[s]) mark[orig[s]] = 1; }
// rank of every live atom's key = its LAMMPS local index; inv[index] = slot; leaver flags re-ordered by index
__global__ void k_orig_rank(int n, const int *alive, int *orig, const int *scan, int *inv, const int *leave, int *leave_by_index)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n || !alive[s]) return;
  const int r = scan[orig[s]];
  orig[s] = r; inv[r] = s; leave_by_index[r] = leave[s];
}
__global__ void k_index_compact(int n, const int *flag, const int *pos, int *list) { int r = blockIdx.x * blockDim.x + threadIdx.x; if (r < n && flag[r]) list[pos[r]] = r; }
// a[0..m): local indices of the leavers, ascending; N: live atoms.  packlist[0..m): slots in the reference's pack order.
// The chain only ever looks at the leavers' own slots and at the last m atoms, so the block first gathers those into three compact arrays
// (work[0..m) slot of leaver k | work[m..2m) slot of the atom at index N-1-j | work[2m..3m) does that atom leave) with parallel loads;
// the one thread that follows the chain then reads sequentially (a dependent random load per link made it ~2 us per leaver).
__global__ void k_holefill(const int *a, int m, const int *nlive, const int *inv, const int *leave, int *orig, int *packlist, int *work)
{
  const int N = *nlive;
  int *slot_a = work, *slot_t = work + m, *leave_t = work + 2 * m;
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    slot_a[k] = inv[a[k]];
    const int idx = N - 1 - k;
    const int t = idx >= 0 ? inv[idx] : 0;
    slot_t[k] = t; leave_t[k] = idx >= 0 ? leave[t] : 0;
  }
  __syncthreads();
  if (threadIdx.x) return;
  int tail = N - 1, np = 0;
  for (int k = 0; k < m && a[k] <= tail; k++) {
    const int i = a[k];
    int cur = slot_a[k];
    for (;;) {
      packlist[np++] = cur;                       // the atom now at index i left: packed
      if (i == tail) { tail--; break; }           // it was the last one: nothing to copy in
      const int j = N - 1 - tail;                 // avec->copy(nlocal-1, i, 1); nlocal--   (at most m atoms are ever taken from the end)
      const int t = slot_t[j]; tail--;
      if (leave_t[j]) { cur = t; continue; }      // the atom copied in leaves as well: index i is examined again
      orig[t] = i;
      break;
    }
  }
}

__global__ void k_pack_exchange(int n, const int *list, CommArrays a, int *alive, double *buf)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int j = list[k];
  double4 x = a.xt[j], vr = a.vr[j], v = a.vm[j], f = a.fd[j], c = a.cgm[j];
  double *b = buf + (size_t)k * NB_EXCHANGE;
  b[0] = x.x; b[1] = x.y; b[2] = x.z; b[3] = (double)tw_type(__double_as_longlong(x.w));
  b[4] = vr.x; b[5] = vr.y; b[6] = vr.z; b[7] = vr.w; b[8] = v.x; b[9] = v.y; b[10] = v.z; b[11] = v.w;
  b[12] = f.x; b[13] = f.y; b[14] = f.z; b[15] = f.w; b[16] = c.x; b[17] = c.y; b[18] = c.z; b[19] = c.w;
  b[20] = a.e[j]; b[21] = a.de[j]; b[22] = a.cv[j]; b[23] = (double)a.tag[j]; b[24] = (double)a.mask[j]; b[25] = 0.0;
  alive[j] = 0;
}
// receiver keeps the atoms that fall inside its own [lo,hi) in `dim` (comm_brick.cpp:657-664)
__global__ void k_unpack_exchange(int n, const double *buf, int dim, double lo, double hi, int first, CommArrays a, int *alive, int *counter, int orig0)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double *b = buf + (size_t)k * NB_EXCHANGE;
  double c = b[dim];
  int i = first + k;                      // every candidate gets a slot; the ones that are not mine stay dead
  bool mine = c >= lo && c < hi;
  alive[i] = mine ? 1 : 0;
  if (!mine) return;
  atomicAdd(counter, 1);
  a.xt[i] = make_double4(b[0], b[1], b[2], __longlong_as_double((long long)pack_tw((int)b[3], 0, 0, 0)));
  a.vr[i] = make_double4(b[4], b[5], b[6], b[7]); a.vm[i] = make_double4(b[8], b[9], b[10], b[11]);
  a.fd[i] = make_double4(b[12], b[13], b[14], b[15]); a.cgm[i] = make_double4(b[16], b[17], b[18], b[19]);
  a.e[i] = b[20]; a.de[i] = b[21]; a.cv[i] = b[22]; a.tag[i] = (int)b[23]; a.mask[i] = (int)b[24];
  a.orig[i] = orig0 + k;                  // arrival order continues this rank's local-index sequence
}
