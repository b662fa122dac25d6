// b200_common.cuh -- shared declarations of the sm_100a SPH engine (libb200sph.so)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/b200_sph.h"

#define MAXT1 8           // per-type tables hold types 1..7
#define MAXTT (MAXT1 * MAXT1)
#define MAXPAIR 16
#define MAXFIX 16
#define WARP 32
#define FULLMASK 0xffffffffu
// neighbor entry: [31] the row particle is the half-list owner of this pair, [30:28] type of j, [27:0] index of j
// (the reference reserves the top 2 bits of its int32 entries the same way, src/lmptype.h:58-59)
#define NBR_OWNER_BIT 0x80000000u
#define NBR_TYPE_SHIFT 28
#define NBR_INDEX_MASK 0x0fffffffu

// force-pass kinds (bit flags of the fused force kernel)
enum { K_TAIT = 1, K_MORRIS = 2, K_TAITMP = 4, K_SURF = 8, K_HEAT = 16, K_HEATMP = 32, K_HEATPC = 64, K_IDEAL = 128, K_LJ = 256 };

// Tables of one pair sub-style, device resident (filled by b200_pair_add).
// Everything indexed [ti * MAXT1 + tj]; cutsq < 0 where the sub-style is not mapped
// onto the type pair, so `rsq < cutsq` is the complete skip-list + cutoff test.
struct PairTab {
  int style, nstep, kind, pad;
  double cutsq[MAXTT];
  double h[MAXTT];
  double c0[MAXTT];     // kernel normalisation folded with powers of 1/h (per style, see fill_tab)
  double c1[MAXTT];
  double visc[MAXTT];   // viscosity | alpha (colorgradient) | D (heat)
  double tc[MAXTT];
  int fixflag[MAXTT];
  int iskip[MAXT1];
  double mass[MAXT1];   // atom->mass (single-phase styles)
  double rho0[MAXT1], B[MAXT1], cs[MAXT1], gamma[MAXT1], rb[MAXT1];
  double self0[MAXT1];  // rhosum self term per type
  int gamma_uniform;    // taitwater/multiphase: all gamma equal -> pressure can be precomputed per particle
  int pad2;
};

struct FixList {
  int n;
  int kind[MAXFIX];     // 1 meso, 2 meso/stationary, 3 gravity, 4 setmeso, 5 enforce2d, 6 setforce (ipar = which components, par = values), 7 setmesode (ipar[1] region kind, par = value, region), 8 addforce (acc = constant components)
  int prog[MAXFIX][3];  // addforce x, y, z / setmeso value: 1 + index of the compiled variable formula (b200_expr.cuh), 0 = constant
  int bit[MAXFIX];
  double acc[MAXFIX][3];
  int ipar[MAXFIX][3];  // setmeso: which, region kind, match_inside
  double par[MAXFIX][7];// setmeso: value, region[6]
};

// geometry handed to kernels by value
struct Geom {
  int dim;
  int periodic[3];
  double boxlo[3], boxhi[3], prd[3], sublo[3], subhi[3];
  double cutghost;
  double slab_lo_hi[3];   // send-left slab upper bound  = sublo + cutghost   (comm_brick.cpp:343)
  double slab_hi_lo[3];   // send-right slab lower bound = subhi - cutghost   (comm_brick.cpp:361)
  // engine cell grid
  double clo[3], cinv[3];
  int nc[3], ncells;
  // the reference's bin grid (Neighbor::setup_bins), for the bit-exact stencil filter
  int nbin[3];
  double bininv[3], binsize[3];
  int sx, sy, sz;
  double cutneighmaxsq;
};

static __host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
static __host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

// ---- packing of xt.w: type (8 bit) | reference bin coords (3 x 18 bit, biased) ----
#define RB_BIAS 4096
static __host__ __device__ __forceinline__ unsigned long long pack_tw(int type, int bx, int by, int bz)
{
  return (unsigned long long)(type & 0xff) | ((unsigned long long)((bx + RB_BIAS) & 0x3ffff) << 8) |
         ((unsigned long long)((by + RB_BIAS) & 0x3ffff) << 26) | ((unsigned long long)((bz + RB_BIAS) & 0x3ffff) << 44);
}
static __host__ __device__ __forceinline__ int tw_type(unsigned long long w) { return (int)(w & 0xff); }
static __host__ __device__ __forceinline__ int tw_bx(unsigned long long w) { return (int)((w >> 8) & 0x3ffff) - RB_BIAS; }
static __host__ __device__ __forceinline__ int tw_by(unsigned long long w) { return (int)((w >> 26) & 0x3ffff) - RB_BIAS; }
static __host__ __device__ __forceinline__ int tw_bz(unsigned long long w) { return (int)((w >> 44) & 0x3ffff) - RB_BIAS; }
