// b200_tile.cuh -- the shared-memory tile path of the pair stages (single-phase and multiphase styles).
//
// Why (profiles/r01_*): with one global gather per neighbor the stage kernels were bound by L1
// wavefronts -- a 32-lane gather touches ~17 distinct 128-byte lines and L1 replays one line per
// ~2 cycles, so 13.4 M gathers cost the whole 1.8 ms of k_force while the fp64 pipe sat at 18 %.
// Here a CTA owns a *tile* = a run of consecutive engine cells of one x-row (<= TILE_ROWS owned
// particles).  Everything its rows can touch lives in the 3x3 (2-D) / 3x3x3 (3-D) cell
// neighbourhood of the run, i.e. in at most 9 contiguous ranges of the cell-sorted particle order:
//   * k_tile_plan   cuts the cell grid into tiles whose candidate set fits in shared memory;
//   * k_tile_build  stages the candidates' positions as fp32 offsets, decides every (row, candidate)
//                   pair by broadcast reads (fp64 only inside a proven error band) and writes each
//                   row's neighbors as 16-bit *slot ids* (position inside the tile's staged set) --
//                   Neighbor::full_bin's list (neigh_full.cpp:241-340), same pair set bit for bit,
//                   half the bytes, in an order that keeps the later LDS.128 reads conflict-free;
//   * k_tile_rhosum / k_tile_force  stage the per-particle records of the tile with TMA bulk
//                   copies (cp.async.bulk + mbarrier), then every lane walks its own row and
//                   reads its neighbors' records from shared memory (LDS.128, ~29 cycles,
//                   no tag lookups) instead of from L1/L2.
// Records are stored as 16-byte parts in separate arrays (P0 = x,y  P1 = z,rho  P2 = vx,vy
// P3 = vz, Tait term  [P4 = e]) so that lanes reading random slots spread over all banks.
//
// Single-phase styles (sph/rhosum, sph/taitwater, sph/taitwater/morris, sph/heatconduction,
// sph/idealgas): every row evaluates its own side of each pair, ghosts included -- ghost x, vest,
// rho, e are fresh copies of their owners (AtomVecMeso::pack_comm, atom_vec_meso.cpp:139-203, and
// the forward_comm_pair of rho, pair_sph_rhosum.cpp:203) and the pair formulas are symmetric, so
// no ghost rows and no reverse communication are needed.
// Multiphase styles (second half of this file): their ghost rho / colorgradient can be one step
// stale in the reference, so entries carry the half-list ownership and a
// ghost flag, tiles of ghost rows accumulate what the reference adds to ghost atoms, and the
// reverse halo stays.
#pragma once
#include "b200_common.cuh"
#include "b200_neigh.cuh"
#include "b200_pair.cuh"

#define TILE_ROWS 256            // owned particles per tile (target; a single denser cell is looped over)
#define TILE_MAXRANGE 9
#define TILE_MAXSEG 18
#define TILE_SLOT_BITS 13
#define TILE_SLOT_MASK 0x1fffu
#define TILE_MAXSLOTS 8190
#ifndef TILE_FORCE_UNROLL
#define TILE_FORCE_UNROLL 8      // neighbors of a group evaluated side by side in the single-phase force body (A/B: tools/gpu_r02aa.sh)
#endif
#ifndef TILE_MPFORCE_UNROLL
#define TILE_MPFORCE_UNROLL 8      // measured on the C3 styles: 8 -> 1.253 ms, 4 -> 1.278, 2 -> 1.332 (single-phase: 8 -> 0.483, 4 -> 0.493)
#endif
constexpr int kForceUnroll = TILE_FORCE_UNROLL, kMpForceUnroll = TILE_MPFORCE_UNROLL;
#define TILE_MP_NPART 9           // record parts of the multiphase force pass (P0..P8, see k_tile_records_mp)
#define TILE_SMEM_MAX 232448     // 227 KB opt-in dynamic shared memory per CTA on sm_100
// multiphase entries carry two more flags (their records are 64-128 B, so a tile never holds more than 2047 slots):
//   [15:13] type of j | [12] the row particle is the reference's "i" of the pair (half-list owner, frozen at build time)
//   | [11] j is a ghost | [10:0] slot
#define TMP_OWNER 0x1000u
#define TMP_GHOST 0x0800u
#define TMP_SLOT_MASK 0x07ffu
#define TMP_MAXSLOTS 2046

struct TileDesc {
  int row0, nrows;               // owned rows [row0, row0 + nrows)
  int c0, ncell;                 // its cells: linear ids c0 .. c0+ncell-1 (one x-row of the engine grid)
  int nrange, nslots, center, ghost;   // ghost: the rows are ghost particles (tile-order indices >= nlocal), candidates owned only
  int rcell[TILE_MAXRANGE];      // range r = one (dy,dz) x-row of candidate cells: first cell (linear id) ...
  int rncell[TILE_MAXRANGE];     // ... number of cells ...
  int rdx[TILE_MAXRANGE];        // ... and x index of its first cell minus x index of c0 (-1 or 0)
  int seg_src[TILE_MAXSEG];      // segment 2r = owned part of range r, 2r+1 = its ghosts: first record (tile order)
  int seg_slot[TILE_MAXSEG + 1]; // first slot of each segment; seg_slot[2*nrange] = nslots
};

// ------------------------------------------------------------------ plan ----
struct TilePlanArgs {
  Geom g; int nlocal, rowcap, slotcap, ghostrows, shrink;   // ghostrows: plan the tiles of the ghost rows (multiphase styles); shrink: see k_tile_plan
  int want, swapdim[3];          // want: -1 all tiles | 0 interior tiles only | 1 boundary tiles only (halo overlap); swapdim[d]: ghosts are exchanged along d
  const int *cso, *csg;
  TileDesc *tiles;
  int *flags;                    // [0] ntiles  [1] max slots  [2] a single cell does not fit  [3] max rows
};

// one warp per x-row of cells: greedy runs of cells while rows <= rowcap and candidates <= slotcap.
// Lane r < 9 owns candidate range r = the x-row at (dy, dz) = (r % 3 - 1, r / 3 - 1); counts are summed over the lanes.
__global__ void k_tile_plan(TilePlanArgs A)
{
  const Geom &g = A.g;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= g.nc[1] * g.nc[2]) return;
  const int cy = w % g.nc[1], cz = w / g.nc[1], base = w * g.nc[0];
  const int *rows = A.ghostrows ? A.csg : A.cso;       // whose rows: owned particles, or the ghosts of the same cells
  const int gr = A.ghostrows;
  const int ny = cy + lane % 3 - 1, nz = cz + lane / 3 - 1;
  const bool rvalid = lane < 9 && ny >= 0 && ny < g.nc[1] && nz >= 0 && nz < g.nc[2];
  const int rbase = rvalid ? (nz * g.nc[1] + ny) * g.nc[0] : 0;
  const unsigned rmask = __ballot_sync(FULLMASK, rvalid);
  const int rrank = __popc(rmask & ((1u << lane) - 1));            // index of this lane's range among the valid ones
  // candidates of this lane's range for the run of cells [x0, x1]: owned, ghost
  auto counts = [&](int x0, int x1, int &no, int &ng) {
    no = ng = 0;
    if (rvalid) {
      int a = rbase + imax(x0 - 1, 0), b = rbase + imin(x1 + 1, g.nc[0] - 1);
      no = A.cso[b + 1] - A.cso[a]; ng = gr ? 0 : A.csg[b + 1] - A.csg[a];
    }
  };
  auto total = [&](int x0, int x1) {
    int no, ng; counts(x0, x1, no, ng);
    int v = no + ng;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
  };
  // Interior tiles (halo overlap): ghosts live in the first / last cell of a swapped dimension and the atoms a halo sends in the
  // first / last two (cell edge >= ghost cutoff), so a tile whose own cells keep to [2, nc-3] there neither reads a ghost nor owns
  // an atom of a send list: it can run while the halo is in flight.
  auto inner1 = [&](int d, int c) { return !A.swapdim[d] || (c >= 2 && c <= g.nc[d] - 3); };
  const bool row_inner = inner1(1, cy) && inner1(2, cz);
  int x0 = 0;
  while (x0 < g.nc[0]) {
    if (rows[base + x0 + 1] == rows[base + x0]) { x0++; continue; }
    int x1 = x0;
    while (x1 + 1 < g.nc[0]) {
      if (A.want >= 0 && inner1(0, x1 + 1) != inner1(0, x0)) break;      // tiles do not straddle the interior / boundary line
      if (rows[base + x1 + 2] - rows[base + x0] > A.rowcap) break;
      if (total(x0, x1 + 1) > A.slotcap) break;
      x1++;
    }
    if (A.shrink)      // (kernels with a run-time lane split) 512 threads serve 256 rows x 2 lanes or 128 rows x 4 lanes: a tile of 129..191 rows wastes more lanes than a shorter one
      while (x1 > x0 && rows[base + x1 + 1] - rows[base + x0] > TILE_ROWS / 2 && rows[base + x1 + 1] - rows[base + x0] < 3 * TILE_ROWS / 4) x1--;
    while (x1 > x0 && rows[base + x1 + 1] == rows[base + x1]) x1--;          // no trailing empty cells
    if (A.want >= 0 && (int)!(row_inner && inner1(0, x0)) != A.want) { x0 = x1 + 1; continue; }
    int no, ng; counts(x0, x1, no, ng);
    int incl = no + ng;                                               // inclusive scan over the lanes -> slot offsets of the ranges
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += v; }
    const int slots = __shfl_sync(FULLMASK, incl, 31), first = incl - (no + ng);
    int t = 0;
    if (lane == 0) {
      if (slots > A.slotcap) atomicExch(&A.flags[2], 1);
      t = atomicAdd(&A.flags[0], 1);
      atomicMax(&A.flags[1], slots); atomicMax(&A.flags[3], rows[base + x1 + 1] - rows[base + x0]);
    }
    t = __shfl_sync(FULLMASK, t, 0);
    TileDesc *d = A.tiles + t;
    if (lane == 0) {
      d->row0 = (gr ? A.nlocal : 0) + rows[base + x0]; d->nrows = rows[base + x1 + 1] - rows[base + x0]; d->c0 = base + x0; d->ncell = x1 - x0 + 1; d->ghost = gr;
      d->nrange = __popc(rmask); d->nslots = slots; d->seg_slot[2 * __popc(rmask)] = slots;
    }
    if (rvalid) {
      const int xa = imax(x0 - 1, 0), xb = imin(x1 + 1, g.nc[0] - 1), a = rbase + xa;
      d->rcell[rrank] = a; d->rncell[rrank] = xb - xa + 1; d->rdx[rrank] = xa - x0;
      d->seg_src[2 * rrank] = A.cso[a]; d->seg_slot[2 * rrank] = first;
      d->seg_src[2 * rrank + 1] = A.nlocal + A.csg[a]; d->seg_slot[2 * rrank + 1] = first + no;
      if (lane == 4) d->center = rrank;                              // (dy, dz) = (0, 0)
    }
    x0 = x1 + 1;
  }
}

// ----------------------------------------------------------------- build ----
struct TileBuildArgs {
  Geom g; int nlocal, ngrp, cap, uni;      // ngrp = groups of 8 entries per row; cap = slots the shared-memory staging holds
  double cutsq_u, farsq_u, midsq_u;        // uni: the one cutneighsq / far / mid threshold of every type pair
  const double4 *xt; const int *gorder; const int *cso, *csg;
  const double *cutneighsq, *farsq, *midsq;
  const TileDesc *tiles; int ntiles; int *counter;
  uint4 *near, *far; int *numneigh, *numfar; int *maxcount;
  int *maxn;                               // longest row (entries), for the counters
  const int *orig; int *rowtile;           // multiphase: LAMMPS local indices (half-list ownership); tile of every owned row (fix phase_change)
};

// 8 entries of a row are one uint4 (16 bits each: [15:13] type of j, [12:0] slot; 0 = empty); group g of row r sits at
// [((r >> 5) * ngrp + g) * 32 + (r & 31)], so a warp of consecutive rows reads one group of each of its rows as 512 contiguous bytes.
// The 8 entries of a group travel through a 128-bit shift register (4 funnel shifts per entry): entry k of a group ends up at
// element k after 8 pushes; finish() shifts a partial group the rest of the way with zeros.
struct RowWriter {
  unsigned x, y, z, w; int n;
  __device__ __forceinline__ RowWriter() : x(0), y(0), z(0), w(0), n(0) {}
  __device__ __forceinline__ void shift(unsigned ent)
  {
    x = __funnelshift_r(x, y, 16); y = __funnelshift_r(y, z, 16); z = __funnelshift_r(z, w, 16); w = __funnelshift_r(w, ent, 16);
  }
  __device__ __forceinline__ void push(unsigned ent, uint4 *base, int ngrp)
  {
    shift(ent);
    if ((n & 7) == 7 && (n >> 3) < ngrp) base[(size_t)(n >> 3) * 32] = make_uint4(x, y, z, w);
    n++;
  }
  __device__ __forceinline__ void finish(uint4 *base, int ngrp)
  {
    if (!(n & 7) || (n >> 3) >= ngrp) return;
    for (int k = n & 7; k < 8; k++) shift(0u);
    base[(size_t)(n >> 3) * 32] = make_uint4(x, y, z, w);
  }
};
// the same from the back of the row: entry k at element 7 - (k & 7) of group ngrp - 1 - (k >> 3)  (the mid zone shares the far array)
struct RowWriterBack {
  unsigned x, y, z, w; int n;
  __device__ __forceinline__ RowWriterBack() : x(0), y(0), z(0), w(0), n(0) {}
  __device__ __forceinline__ void shift(unsigned ent)
  {
    w = __funnelshift_l(z, w, 16); z = __funnelshift_l(y, z, 16); y = __funnelshift_l(x, y, 16); x = (x << 16) | ent;
  }
  __device__ __forceinline__ void push(unsigned ent, uint4 *base, int ngrp)
  {
    shift(ent);
    const int g = ngrp - 1 - (n >> 3);
    if ((n & 7) == 7 && g >= 0) base[(size_t)g * 32] = make_uint4(x, y, z, w);
    n++;
  }
  __device__ __forceinline__ void finish(uint4 *base, int ngrp)
  {
    const int g = ngrp - 1 - (n >> 3);
    if (!(n & 7) || g < 0) return;
    for (int k = n & 7; k < 8; k++) shift(0u);
    base[(size_t)g * 32] = make_uint4(x, y, z, w);
  }
};

// Near rows are written in a bank-aware order.  The stage kernels read a neighbor's record parts with 128-bit LDS, served per
// quarter-warp (8 lanes x 16 B): conflict-free iff the 8 slots differ mod 8.  The order of a row's entries is free, so the k-th
// entry of residue class c = slot mod 8 of lane q (= row & 7 inside its tile) goes to position 8 k + ((c - q) mod 8): at step t
// the 8 lanes of a quarter-warp then read 8 different bank groups.  Classes are not equally full; finish() moves the entries
// that stick out beyond ceil(n / 8) groups into the holes of the shorter classes (~10 % of a row, the only possible conflicts)
// and zeroes what stays empty.  (profiles/r01_tile_*: in natural order conflicts were 54 % of k_tile_force's smem wavefronts.)
struct NearWriter {
  unsigned clo, chi; int n;               // 8 x 8-bit class counters (classes 0-3 | 4-7)
  __device__ __forceinline__ NearWriter() : clo(0), chi(0), n(0) {}
  __device__ __forceinline__ int count(int c) const { return (int)(((c & 4) ? chi : clo) >> ((c & 3) * 8)) & 0xff; }
  // position 8 k + ((c - q) mod 8) = group k, element (c - q) mod 8; a group is 32 rows x 8 entries = 256 shorts apart
  __device__ __forceinline__ static int off_of(int c, int k, int q) { return k * 256 + ((c - q) & 7); }
  __device__ __forceinline__ void push(unsigned ent, int q, unsigned short *base, int stride)
  {
    const int c = ent & 7, sh = (c & 3) * 8;
    const bool hi = (c & 4) != 0;
    const int k = (int)((hi ? chi : clo) >> sh) & 0xff;
    const unsigned inc = (unsigned)(k < 255) << sh;
    if (hi) chi += inc; else clo += inc;
    if (8 * k < stride) base[off_of(c, k, q)] = (unsigned short)ent;
    n++;
  }
  // returns the extent the row needed (in entries); the row is valid iff that is <= stride
  __device__ __forceinline__ int finish(int q, unsigned short *base, int stride)
  {
    int mx = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) mx = max(mx, count(c));
    if (mx >= 255) return 1 << 20;
    if (8 * mx > stride) return 8 * mx;
    const int D = (n + 7) >> 3;
    int ch = 0, kh = count(0);
    for (int c = 0; c < 8; c++)
      for (int k = D; k < count(c); k++) {
        while (kh >= D) { ch++; kh = count(ch); }
        base[off_of(ch, kh, q)] = base[off_of(c, k, q)];
        kh++;
      }
    for (;;) {
      while (kh >= D) { if (++ch == 8) return 8 * mx; kh = count(ch); }
      base[off_of(ch, kh, q)] = 0;
      kh++;
    }
  }
};

__device__ __forceinline__ int tile_slot_src(const TileDesc &D, int slot, int nlocal, const int *gorder)
{
  int s = 0;
  while (slot >= D.seg_slot[s + 1]) s++;
  int src = D.seg_src[s] + (slot - D.seg_slot[s]);
  return (s & 1) ? nlocal + gorder[src - nlocal] : src;
}

// the exact pair test of Neighbor::full_bin (neigh_full.cpp:241-340) as k_build restates it: 0 = not a neighbor, 1 = near row, 2 = far row
template <bool UNI> __device__ __forceinline__ int tile_exact_class(const TileBuildArgs &A, const double4 pi, const double4 pj)
{
  const Geom &g = A.g;
  const unsigned long long wi = (unsigned long long)__double_as_longlong(pi.w), wj = (unsigned long long)__double_as_longlong(pj.w);
  const double rsq = rsq_nofma(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
  const int tij = tw_type(wi) * MAXT1 + tw_type(wj);
  if (!(rsq <= (UNI ? A.cutsq_u : A.cutneighsq[tij]))) return 0;       // UNI: the one threshold of every type pair travels as a kernel parameter
  const double cutmaxsq = g.cutneighmaxsq;
  if (rsq >= cutmaxsq * (1.0 - 1.0e-9)) {          // the reference's own bin stencil (neigh_stencil.cpp:434-448), see k_build
    int dbx = abs(tw_bx(wj) - tw_bx(wi)), dby = abs(tw_by(wj) - tw_by(wi)), dbz = abs(tw_bz(wj) - tw_bz(wi));
    if (dbx > g.sx || dby > g.sy || dbz > g.sz) return 0;
    double ex = dbx ? (dbx - 1) * g.binsize[0] : 0.0, ey = dby ? (dby - 1) * g.binsize[1] : 0.0, ez = dbz ? (dbz - 1) * g.binsize[2] : 0.0;
    if (!(rsq_nofma(ex, ey, ez) < cutmaxsq)) return 0;
  }
  return rsq >= (UNI ? A.farsq_u : A.farsq[tij]) ? 2 : 1;
}

// One CTA per tile.  Candidate positions are staged as fp32 offsets from the tile's corner; every (row, candidate) pair is decided in
// fp32 against thresholds widened by a proven error band (|rsq32 - rsq| <= 2^-23 (2 sqrt(3) r (2E + r) + 4 r^2), E = largest offset),
// and only pairs inside the band -- a ~1e-5 fraction -- take the exact fp64 test, so the list is bit-for-bit the one the fp64 test gives.
// Phase A: 32 candidates x 3 compares -> bit masks (broadcast float4 reads).  Phase B: entries straight from the masks.
// MP (multiphase styles): entries also carry the half-list ownership of the pair (neigh_derive.cpp:83-145: local index order
// among owned atoms, the "above/right" rule for a ghost) and a ghost flag; tiles of ghost rows list, for a ghost g, the owned
// atoms whose half list holds (i,g) -- what the reference adds to ghost atoms and reverse-communicates.
// NT: 256 threads for big tiles (C2: 8 chunks of 32 rows per tile), 128 when the tiles are small (more CTAs per SM to overlap the
// per-tile barriers); the (cell, chunk) work items of a tile are handed to the warps through a shared counter.
// ZONES = false (decks with skin 0, e.g. the shipped multiphase decks: rebuilt every step, every entry is a near entry): the far / mid
// compares of phase A and their entry loops are compiled out -- 2 of the 4 compares per candidate.
template <bool UNI, bool MP, int NT, bool ZONES>
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 0) k_tile_build(const __grid_constant__ TileBuildArgs A)
{
  constexpr int TILE_BUILD_NT = NT;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  const int cap4 = ((A.cap + 3) & ~3) + 4;
  float *fx = (float *)tile_smem, *fy = fx + cap4, *fz = fy + cap4;
  int *so = (int *)(fz + cap4);                                // MP: LAMMPS local index of the owned candidates
  unsigned char *ty = (unsigned char *)(MP ? (void *)(so + cap4) : (void *)so);
  __shared__ TileDesc D;
  __shared__ int s_tile, s_item;
  __shared__ unsigned s_emax;
  __shared__ float s_thr[MAXTT][6];                            // non-uniform cutoffs: far_lo, far_hi, cut_lo, cut_hi, mid_lo, mid_hi per type pair
  const int tid = threadIdx.x, lane = tid & 31;
  const Geom &g = A.g;
  const int ntiles = A.ntiles;
  for (;;) {
    __syncthreads();
    if (tid == 0) { s_tile = atomicAdd(A.counter, 1); s_emax = 0; s_item = 0; }
    __syncthreads();
    const int t = s_tile;
    if (t >= ntiles) break;
    for (int k = tid; k < (int)(sizeof(TileDesc) / 4); k += TILE_BUILD_NT) ((int *)&D)[k] = ((const int *)(A.tiles + t))[k];
    __syncthreads();
    const bool gt = MP && D.ghost;
    // corner of the tile's candidate region (cell c0 shifted by one cell in every direction)
    const int cx0 = D.c0 % g.nc[0], cy0 = (D.c0 / g.nc[0]) % g.nc[1], cz0 = D.c0 / (g.nc[0] * g.nc[1]);
    const double ox = g.clo[0] + (cx0 - 1) / g.cinv[0], oy = g.clo[1] + (cy0 - 1) / g.cinv[1], oz = g.clo[2] + (cz0 - 1) / g.cinv[2];
    float emax = 0.f;
    for (int s = 0; s < 2 * D.nrange; s++) {
      int s0 = D.seg_slot[s], n = D.seg_slot[s + 1] - s0, src0 = D.seg_src[s];
      for (int k = tid; k < n; k += TILE_BUILD_NT) {
        int src = src0 + k;
        if (s & 1) src = A.nlocal + A.gorder[src - A.nlocal];
        double4 p = A.xt[src];
        float x = (float)(p.x - ox), y = (float)(p.y - oy), z = (float)(p.z - oz);
        fx[s0 + k] = x; fy[s0 + k] = y; fz[s0 + k] = z; ty[s0 + k] = (unsigned char)tw_type(__double_as_longlong(p.w));
        if (MP) so[s0 + k] = (s & 1) ? 0 : A.orig[src];
        emax = fmaxf(emax, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
      }
    }
    if (gt)                                                    // ghost rows are not among the candidates: their offsets count too
      for (int k = tid; k < D.nrows; k += TILE_BUILD_NT) {
        double4 p = A.xt[A.nlocal + A.gorder[D.row0 + k - A.nlocal]];
        emax = fmaxf(emax, fmaxf(fabsf((float)(p.x - ox)), fmaxf(fabsf((float)(p.y - oy)), fabsf((float)(p.z - oz)))));
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(FULLMASK, emax, o));
    if (lane == 0) atomicMax(&s_emax, __float_as_uint(emax));
    __syncthreads();
    const double E = (double)__uint_as_float(s_emax) * 1.0001;
    const float ztol = (float)(E * 2.4e-7);                   // two fp32 roundings of an offset <= E, doubled
    // thresholds with the error band: a pair is "sure inside" below lo, "sure outside" at or above hi
    auto band = [&](double thr, float &lo, float &hi) {
      double r = sqrt(fmax(thr, 0.0));
      double err = 2.4e-7 * (3.4642 * r * (2.0 * E + r) + 4.0 * r * r) + 1e-37;      // 2^-22: twice the bound
      lo = __double2float_rd(thr - err); hi = __double2float_ru(thr + err);
    };
    float far_lo, far_hi, cut_lo, cut_hi, mid_lo, mid_hi;
    if (UNI) { band(A.farsq_u, far_lo, far_hi); band(A.cutsq_u, cut_lo, cut_hi); band(A.midsq_u, mid_lo, mid_hi); }
    else {
      band(g.cutneighmaxsq, cut_lo, cut_hi); far_lo = far_hi = mid_lo = mid_hi = 0.f;
      for (int k = tid; k < MAXTT; k += TILE_BUILD_NT) {
        band(fmin(A.farsq[k], 1e30), s_thr[k][0], s_thr[k][1]); band(A.cutneighsq[k], s_thr[k][2], s_thr[k][3]); band(fmin(A.midsq[k], 1e30), s_thr[k][4], s_thr[k][5]);
      }
      __syncthreads();
    }
    // work items: (cell of the tile, chunk of 32 of its rows); item -> warp round robin
    const int *rowstart = gt ? A.csg : A.cso;
    constexpr bool DYN = NT < 256;                               // big tiles: one chunk per warp and round, static is cheaper
    int item = 0, mine = tid >> 5;
    if (DYN) { if (lane == 0) mine = atomicAdd(&s_item, 1); mine = __shfl_sync(FULLMASK, mine, 0); }
    for (int ci = 0; ci < D.ncell; ci++) {
      const int cr0 = (gt ? A.nlocal : 0) + rowstart[D.c0 + ci], cnr = rowstart[D.c0 + ci + 1] - rowstart[D.c0 + ci];
      for (int rb = 0; rb < cnr; rb += 32, item++) {
        if (item != mine) continue;
        const bool valid = rb + lane < cnr;
        const int row = cr0 + rb + lane;                                      // tile-order index
        const int dev = !valid ? 0 : (gt ? A.nlocal + A.gorder[row - A.nlocal] : row);   // device index
        const int myslot = (valid && !gt) ? D.seg_slot[2 * D.center] + (row - D.seg_src[2 * D.center]) : -1;
        float xi = 1e30f, yi = 1e30f, zi = 1e30f; int ti = 0, oi = 0;
        double4 pI = make_double4(0, 0, 0, 0);
        if (valid && (MP || gt)) pI = A.xt[dev];
        if (valid) {
          if (!gt) { xi = fx[myslot]; yi = fy[myslot]; zi = fz[myslot]; ti = ty[myslot]; if (MP) oi = so[myslot]; }
          else { xi = (float)(pI.x - ox); yi = (float)(pI.y - oy); zi = (float)(pI.z - oz); ti = tw_type(__double_as_longlong(pI.w)); }
        }
        const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
        uint4 *nrow = A.near + rbase, *frow = A.far + rbase;
        NearWriter wn; RowWriter wf; RowWriterBack wm;
        unsigned short *nrow16 = (unsigned short *)nrow;
        const int q = (row - D.row0) & 7, stride = A.ngrp * 8;

        // candidates in slots [s0, s1) of the segment that starts at slot slot0 / record src0 (jghost: a ghost segment)
        auto interval = [&](int s0, int s1, int slot0, int src0, bool jghost) {
          auto dev_of = [&](int slot) { int src = src0 + (slot - slot0); return jghost ? A.nlocal + A.gorder[src - A.nlocal] : src; };
          // "is b above/right of a" (neigh_derive.cpp:121-134) from the staged fp32 offsets when z differs by more than their
          // rounding error, else from the fp64 coordinates (ties in z, then y, then x must be exact)
          auto above = [&](bool row_is_a, int slot) {
            const float za = row_is_a ? zi : fz[slot], zb = row_is_a ? fz[slot] : zi;
            if (zb > za + ztol) return true;
            if (zb < za - ztol) return false;
            const double4 pj = A.xt[dev_of(slot)];
            return row_is_a ? ghost_above(pI.x, pI.y, pI.z, pj.x, pj.y, pj.z) : ghost_above(pj.x, pj.y, pj.z, pI.x, pI.y, pI.z);
          };
          // multiphase: ownership / ghost flags of an entry; false = the pair does not belong in this row
          auto flags = [&](int slot, unsigned &ent) {
            if (!MP) return true;
            if (gt) {                                           // (owned j, ghost row i): kept by j's half list iff i is above/right of j
              if (!above(false, slot)) return false;
              ent |= TMP_OWNER;
            } else if (jghost) {
              if (above(true, slot)) ent |= TMP_OWNER;
              ent |= TMP_GHOST;
            } else if (oi < so[slot]) ent |= TMP_OWNER;
            return true;
          };
          for (int bj = s0 & ~3; bj < s1; bj += 32) {
            // phase A: 32 candidates, three compares each.  in: surely inside the cutoff; mb: inside or in its error band;
            // fr: surely in the far zone (one-sided: an entry just beyond the far threshold may stay in the near row, where it is only tested more often)
            // the sign bit of (rsq - threshold) is the compare; a funnel shift appends it to the mask (2 instructions per compare)
            unsigned in = 0, mb = 0, fr = 0, md = 0; int nq = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const float4 X = *(const float4 *)(fx + bj + 4 * k), Y = *(const float4 *)(fy + bj + 4 * k), Z = *(const float4 *)(fz + bj + 4 * k);
              const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
              for (int c = 0; c < 4; c++) {
                const float dx = xi - xs[c], dy = yi - ys[c], dz = zi - zs[c];
                const float rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                in = __funnelshift_l(__float_as_uint(rsq - cut_lo), in, 1);
                mb = __funnelshift_l(__float_as_uint(rsq - cut_hi), mb, 1);
                if (UNI && ZONES) { fr = __funnelshift_l(__float_as_uint(rsq - far_hi), fr, 1); md = __funnelshift_l(__float_as_uint(rsq - mid_hi), md, 1); }   // 1 = NOT surely far / mid
              }
              nq = k + 1;
              if (bj + 4 * k + 4 >= s1) break;
            }
            // candidate bj + i sits at bit 4 nq - 1 - i: back to natural order
            in = __brev(in << (32 - 4 * nq)); mb = __brev(mb << (32 - 4 * nq));
            if (UNI && ZONES) { fr = ~__brev(fr << (32 - 4 * nq)); md = ~__brev(md << (32 - 4 * nq)); }
            else if (UNI) fr = md = 0;
            // only the slots of [s0, s1), and never the row particle itself
            unsigned vm = (s1 - bj >= 32) ? 0xffffffffu : ((1u << (s1 - bj)) - 1u);
            if (bj < s0) vm &= ~((1u << (s0 - bj)) - 1u);
            if (myslot >= bj && myslot < bj + 32) vm &= ~(1u << (myslot - bj));
            if (!valid) vm = 0;
            in &= vm; mb &= vm;
            unsigned border = UNI ? (mb ^ in) : mb;                  // per-type thresholds: classify every coarse hit
            if (!UNI) in = 0;
            if (UNI && border) {
              // band pairs take the fp64 test on the particles' own coordinates (global memory).  On a lattice whole shells of
              // neighbors sit exactly on the cutoff and land here (profiles/r02_mpbuild_lines.txt: 28 % of the samples of the C3
              // build, all of them waiting for these loads), so the next pair's coordinates are fetched while this one is decided.
              if (!(MP || gt)) pI = A.xt[dev];
              int idx = __ffs((int)border) - 1; border &= border - 1;
              double4 pj = A.xt[dev_of(bj + idx)];
              for (;;) {
                const int cur = idx; const double4 pc = pj;
                const bool more = border != 0;
                if (more) { idx = __ffs((int)border) - 1; border &= border - 1; pj = A.xt[dev_of(bj + idx)]; }
                const int cls = tile_exact_class<true>(A, pI, pc);
                if (cls) in |= 1u << cur;
                if (ZONES) {
                  if (cls == 2) fr |= 1u << cur; else fr &= ~(1u << cur);
                  if (cls == 3) md |= 1u << cur; else md &= ~(1u << cur);
                }
                if (!more) break;
              }
            }
            while (!UNI && border) {
              const int idx = __ffs((int)border) - 1; border &= border - 1;
              const int slot = bj + idx;
              int cls = -1;
              {
                const float dx = xi - fx[slot], dy = yi - fy[slot], dz = zi - fz[slot];
                const float rsq = dx * dx + dy * dy + dz * dz;
                const float *th = s_thr[ti * MAXT1 + ty[slot]];
                if (rsq >= th[3]) cls = 0;                         // surely outside the neighbor cutoff
                else if (rsq < th[2]) cls = rsq >= th[1] ? 2 : (rsq >= th[5] ? 3 : 1);    // surely inside: far / mid zone only if surely beyond that threshold
              }
              if (cls < 0) {
                if (!(MP || gt)) pI = A.xt[dev];
                cls = tile_exact_class<false>(A, pI, A.xt[dev_of(slot)]);
              }
              if (cls) in |= 1u << idx;
              if (cls == 2) fr |= 1u << idx; else fr &= ~(1u << idx);
              if (cls == 3) md |= 1u << idx; else md &= ~(1u << idx);
            }
            // phase B: entries straight from the masks (one-sided zone thresholds: md = surely beyond cut + mid margin, fr = surely beyond cut + far margin)
            unsigned nearm = ZONES ? in & ~fr & ~md : in, midm = ZONES ? in & ~fr & md : 0u, farm = ZONES ? in & fr : 0u;
            while (nearm) {
              const int idx = __ffs((int)nearm) - 1; nearm &= nearm - 1;
              const int slot = bj + idx;
              unsigned ent = ((unsigned)ty[slot] << TILE_SLOT_BITS) | (unsigned)slot;
              if (flags(slot, ent)) wn.push(ent, q, nrow16, stride);
            }
            while (midm) {
              const int idx = __ffs((int)midm) - 1; midm &= midm - 1;
              const int slot = bj + idx;
              unsigned ent = ((unsigned)ty[slot] << TILE_SLOT_BITS) | (unsigned)slot;
              if (flags(slot, ent)) wm.push(ent, frow, A.ngrp);
            }
            while (farm) {
              const int idx = __ffs((int)farm) - 1; farm &= farm - 1;
              const int slot = bj + idx;
              unsigned ent = ((unsigned)ty[slot] << TILE_SLOT_BITS) | (unsigned)slot;
              if (flags(slot, ent)) wf.push(ent, frow, A.ngrp);
            }
          }
        };

        for (int r = 0; r < D.nrange; r++) {
          int k0 = imax(ci - 1 - D.rdx[r], 0), k1 = imin(ci + 1 - D.rdx[r], D.rncell[r] - 1);
          if (k0 > k1) continue;
          int ca = D.rcell[r] + k0, cb = D.rcell[r] + k1, cr = D.rcell[r];
          interval(D.seg_slot[2 * r] + A.cso[ca] - A.cso[cr], D.seg_slot[2 * r] + A.cso[cb + 1] - A.cso[cr], D.seg_slot[2 * r], D.seg_src[2 * r], false);
          if (gt) continue;
          int ga = A.csg[ca] - A.csg[cr], gb = A.csg[cb + 1] - A.csg[cr];
          if (gb > ga) interval(D.seg_slot[2 * r + 1] + ga, D.seg_slot[2 * r + 1] + gb, D.seg_slot[2 * r + 1], D.seg_src[2 * r + 1], true);
        }
        if (valid) {
          const int ext = wn.finish(q, nrow16, stride);
          wf.finish(frow, A.ngrp); wm.finish(frow, A.ngrp);
          A.numneigh[row] = wn.n; A.numfar[row] = wf.n | (wm.n << 16);
          if (A.rowtile && !gt) A.rowtile[row] = t;
          atomicMax(A.maxcount, max(ext, ((wf.n + 7) & ~7) + ((wm.n + 7) & ~7)));
          atomicMax(A.maxn, wn.n + wf.n + wm.n);     // far groups from the front and mid groups from the back must not meet
        }
        if (DYN) { if (lane == 0) mine = atomicAdd(&s_item, 1); mine = __shfl_sync(FULLMASK, mine, 0); }
        else mine += NT / 32;
      }
    }
  }
}

// slot ids -> device particle indices, row-major [row][width] (tests / b200_get_neighbor_list only)
struct TileExportArgs {
  int nlocal, ngrp, width, slot_mask;
  const int *gorder; const TileDesc *tiles; int ntiles;
  const uint4 *near, *far; const int *numneigh, *numfar;
  int *out;
};
__global__ void k_tile_export(TileExportArgs A)
{
  __shared__ TileDesc D;
  const int ntiles = A.ntiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads();
    for (int k = threadIdx.x; k < (int)(sizeof(TileDesc) / 4); k += blockDim.x) ((int *)&D)[k] = ((const int *)(A.tiles + t))[k];
    __syncthreads();
    for (int rt = threadIdx.x; rt < D.nrows; rt += blockDim.x) {
      int row = D.row0 + rt, nn = A.numneigh[row], nf = A.numfar[row], o = 0;
      const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
      for (int pass = 0; pass < 3; pass++) {           // near | far | mid (from the back of the far row)
        const unsigned short *p = (const unsigned short *)((pass ? A.far : A.near) + rbase);
        int n = ((pass == 0 ? nn : (pass == 1 ? (nf & 0xffff) : (nf >> 16))) + 7) & ~7;
        for (int k = 0; k < n; k++) {
          int ent = pass == 2 ? p[(size_t)(A.ngrp - 1 - (k >> 3)) * 32 * 8 + (k & 7)] : p[(size_t)(k >> 3) * 32 * 8 + (k & 7)];
          if (!ent) continue;
          int slot = ent & A.slot_mask;
          int s = 0;
          while (slot >= D.seg_slot[s + 1]) s++;
          int src = D.seg_src[s] + (slot - D.seg_slot[s]);
          if (s & 1) src = A.nlocal + A.gorder[src - A.nlocal];
          if (o < A.width) A.out[(size_t)row * A.width + o] = src;
          o++;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ zones ----
// cell of every owned row (rows are the owned atoms in cell order; perm2 = row -> slot before the sort)
__global__ void k_row_cells(int n, const int *perm2, const int *cellid, int *rowcell)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rowcell[i] = cellid[perm2[i]];
}
// A tile's mid / far rows are due once 2 * dmax reaches their margin, dmax = largest displacement since the build over the tile's
// rows AND candidates, i.e. over the atoms of the cells of its candidate ranges (celld).  A handful of fast atoms (a jet, a free
// surface) then no longer switches the far rows on for the whole domain.
__global__ void k_tile_zone(const TileDesc *tiles, int ntiles, const unsigned *celld, double marginsq, double midmarginsq, unsigned char *tzone)
{ // one warp per tile: the lanes share out the cells of its candidate ranges (a thread per tile walked ~50 dependent loads: 19 us per launch)
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= ntiles) return;
  const TileDesc &D = tiles[t];
  unsigned m = 0;
  for (int r = 0; r < D.nrange; r++)
    for (int k = lane; k < D.rncell[r]; k += 32) m = max(m, celld[D.rcell[r] + k]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(FULLMASK, m, o));
  const double d = 4.0 * (double)__uint_as_float(m);
  if (lane == 0) tzone[t] = (unsigned char)((d >= 0.99 * marginsq ? 1 : 0) | (d >= 0.99 * midmarginsq ? 2 : 0));
}

// --------------------------------------------------------------- records ----
// Per-pass records in tile order (owned atoms in device order, then the ghosts in cell order = gorder), one
// double2 array per part: P0 = x,y   P1 = z,rho   [P2 = vest.x,vest.y   P3 = vest.z, Tait term]   [Pe = e,0]
struct TileRecArgs {
  int nlocal, nall, pstride, force, epart;   // epart < 0: no energy part
  int i0, i1;                                // records [i0, i1) of the tile order (owned first, then ghosts)
  const int *gorder; const double4 *xt, *vr; const double *e; const PairTab *fluid;
  double2 *rec;
};
__global__ void k_tile_records(TileRecArgs A)
{
  int i = A.i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.i1) return;
  int src = i < A.nlocal ? i : A.nlocal + A.gorder[i - A.nlocal];
  double4 x = A.xt[src], v = A.vr[src];
  A.rec[i] = make_double2(x.x, x.y);
  A.rec[(size_t)A.pstride + i] = make_double2(x.z, v.w);
  if (A.force) {
    double pf = 0.0;
    if (A.fluid) {
      int t = tw_type(__double_as_longlong(x.w));
      if (A.fluid->style == B200_PAIR_IDEALGAS) pf = 0.4 * A.e[src] / A.fluid->mass[t] / v.w;     // p / rho^2, pair_sph_idealgas.cpp:94
      else {                           // B((rho/rho0)^7 - 1)/rho^2, pair_sph_taitwater.cpp:118-120
        double tmp = v.w / A.fluid->rho0[t], fi = tmp * tmp * tmp;
        pf = A.fluid->B[t] * (fi * fi * tmp - 1.0) / (v.w * v.w);
      }
    }
    A.rec[(size_t)2 * A.pstride + i] = make_double2(v.x, v.y);
    A.rec[(size_t)3 * A.pstride + i] = make_double2(v.z, pf);
  }
  if (A.epart >= 0) A.rec[(size_t)A.epart * A.pstride + i] = make_double2(A.e[src], 0.0);
}

// ---------------------------------------------------- TMA / mbarrier glue ---
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{ // 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (UBLKCP in SASS)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint4 ldg_nc_u4(const uint4 *p)
{
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Per-table constants of a sub-style whose coefficients are the same for every mapped type pair (the common deck:
// `pair_coeff * *`): they travel as kernel parameters (constant bank -> registers) instead of shared-memory table reads.
struct TileUni {
  unsigned long long mapmask;   // bit ti*8+tj: the sub-style acts on the pair (type 0 = empty entry is never mapped)
  double cutsq, h, c0, c1, visc, mass, cs, self;
};

struct TileArgs {
  int nlocal, ngrp, pstride, cap;            // cap = slots per part in shared memory
  const double2 *rec;
  const uint4 *near, *far; const int *numneigh, *numfar; const int *scan_far;
  const TileDesc *tiles; int ntiles; int *counter;      // tiles[0 .. ntiles): the launch's share of the plan
  const double4 *xt;
  double4 *vr_out, *fd; double *de;
  const PairTab *tab[3];
  TileUni uni[3];
  // multiphase styles
  const double4 *vm; double4 *cg_out; const int *gorder; int dim;
  double *virow;                             // VIR kernels: per-row virial sums [row][6]
  const unsigned char *tzone;                // single-phase stage kernels: per-tile zone flags (bit 0 far, bit 1 mid), or NULL -> scan_far
  int accum;                                 // force kernels: 0 = first force pass since force_clear (f, drho, de are zero: plain stores), 1 = add to what is there
};

// shared-memory map of the stage kernels: [NPARTS][cap] double2 | PairTab[NK] | TileDesc[2] | mbarrier | tile id[2]
template <int NPARTS, int NK> struct TileSmem {
  double2 *part; PairTab *T; TileDesc *D; unsigned long long *bar; int *tile;
  __device__ __forceinline__ TileSmem(unsigned char *base, int cap)
  {
    part = (double2 *)base;
    T = (PairTab *)(base + (size_t)NPARTS * cap * 16);
    D = (TileDesc *)(T + NK);
    bar = (unsigned long long *)(D + 2);
    tile = (int *)(bar + 1);
  }
  static size_t bytes(int cap) { return (size_t)NPARTS * cap * 16 + NK * sizeof(PairTab) + 2 * sizeof(TileDesc) + 16; }
};

// The tile loop of a persistent stage CTA.  One barrier per tile: while tile t is evaluated, warp 1 draws the next tile id
// and copies its descriptor into the other descriptor slot; release() (all rows of t done) lets warp 0 issue the bulk copies
// of tile t+1 at once.  PM = record parts to stage (bit p = part p of A.rec), packed densely in shared memory.
template <int PM, int NK, int NT> struct TileLoop {
  static constexpr int NP = __builtin_popcount(PM);
  const TileArgs &A; TileSmem<NP, NK> &S; int ntiles, cur; unsigned phase;
  __device__ __forceinline__ TileLoop(const TileArgs &A_, TileSmem<NP, NK> &S_, int ntiles_) : A(A_), S(S_), ntiles(ntiles_), cur(0), phase(0) {}
  __device__ __forceinline__ void issue(const TileDesc &D)      // warp 0
  {
    const int lane = threadIdx.x;
    if (lane == 0) mbar_expect_tx(S.bar, (unsigned)D.nslots * 16u * NP);
    __syncwarp();
    for (int s = lane; s < 2 * D.nrange; s += 32) {
      int s0 = D.seg_slot[s], n = D.seg_slot[s + 1] - s0;
      if (n > 0) {
        int q = 0;
#pragma unroll
        for (int p = 0; p < TILE_MP_NPART; p++)
          if (PM & (1 << p)) { bulk_g2s(S.part + (size_t)q * A.cap + s0, A.rec + (size_t)p * A.pstride + D.seg_src[s], (unsigned)n * 16u, S.bar); q++; }
      }
    }
  }
  __device__ __forceinline__ void fetch(int slot, int lane)       // one warp: next tile id + descriptor -> slot
  {
    int t = 0;
    if (lane == 0) { t = atomicAdd(A.counter, 1); S.tile[slot] = t; }
    t = __shfl_sync(FULLMASK, t, 0);
    if (t < ntiles) {
      for (int k = lane; k < (int)(sizeof(TileDesc) / 4); k += 32) ((int *)(S.D + slot))[k] = ((const int *)(A.tiles + t))[k];
    }
  }
  __device__ __forceinline__ void start()
  {
    if (threadIdx.x < 32) fetch(0, threadIdx.x);
    __syncthreads();                                     // also: tables loaded, mbarrier initialised
    if (S.tile[0] < ntiles && threadIdx.x < 32) issue(S.D[0]);
  }
  // the descriptor of the current tile (visible since the last barrier), or nullptr when the tiles are used up.  Its records may
  // still be in flight: the caller issues the global loads of its rows (counts, first entries) and only then calls wait(), so
  // that their latency overlaps the bulk copies instead of following them (profiles/r02_force_*: long-scoreboard stalls)
  __device__ __forceinline__ const TileDesc *peek()
  {
    if (S.tile[cur] >= ntiles) return nullptr;
    if ((threadIdx.x >> 5) == 1) fetch(cur ^ 1, threadIdx.x & 31);
    return S.D + cur;
  }
  __device__ __forceinline__ void wait() { mbar_wait(S.bar, phase); phase ^= 1; }
  __device__ __forceinline__ const TileDesc *acquire() { const TileDesc *d = peek(); if (d) wait(); return d; }
  __device__ __forceinline__ void release()
  {
    __syncthreads();                                     // everyone is done with this tile's records; the next descriptor is visible
    cur ^= 1;
    if (S.tile[cur] < ntiles && threadIdx.x < 32) issue(S.D[cur]);
  }
};

// branch-free fp64 sqrt and division: MUFU seed (~2^-22), one coupled Newton step (~2^-43), one residual correction
// (error ~ the square of that, i.e. below 1 ulp).  The CUDA built-ins carry a slow-path branch that splits the pair body
// into basic blocks and keeps ptxas from interleaving the 8 unrolled neighbors of a group.
__device__ __forceinline__ double fast_sqrt(double a)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, hh = 0.5 * y;
  const double e = fma(-hh, g, 0.5);
  g = fma(g, e, g); hh = fma(hh, e, hh);
  return fma(fma(-g, g, a), hh, g);
}
__device__ __forceinline__ double fast_div(double n, double d)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  const double q = n * y;
  return fma(fma(-d, q, n), y, q);
}
// the same without the last correction: relative error <= ~1e-13 (seed 2^-22 squared, times 1.5) resp. one rounding of the
// reciprocal -- three orders inside the 1e-10 per-step budget; used in the single-phase force body, where the fp64 pipe binds
__device__ __forceinline__ double fast_sqrt13(double a)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = a * y, hh = 0.5 * y;
  return fma(g, fma(-hh, g, 0.5), g);
}
__device__ __forceinline__ double fast_div15(double n, double d)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  y = fma(y, fma(-d, y, 1.0), y);
  y = fma(y, fma(-d, y, 1.0), y);
  return n * y;
}
// the cutoff tests of the stage kernels may contract to FMA: every kernel and its derivative vanish at the cutoff, so a
// pair within an ulp of it contributes nothing either way (the neighbor *lists* are decided without FMA, see k_tile_build)
__device__ __forceinline__ double rsq_fma(double dx, double dy, double dz) { return fma(dz, dz, fma(dy, dy, dx * dx)); }
__device__ __forceinline__ bool dpos(double a) { return (__double2hiint(a) | __double2loint(a)) != 0; }   // a > 0 for a >= 0, on the integer pipe

// ---------------------------------------------------------------- density ---
// PairSPHRhoSum::compute, pair_sph_rhosum.cpp:112-197 (full list; quadric kernel, per-type mass)
// (SPLIT <= 2: 64 registers, so that two CTAs share an SM when their tiles fit -- one stages while the other computes)
template <int SPLIT, bool UNI>
__global__ void __launch_bounds__(TILE_ROWS * SPLIT, SPLIT <= 2 ? 2 : 1) k_tile_rhosum(const __grid_constant__ TileArgs A)
{
  constexpr int NT = TILE_ROWS * SPLIT, LPW = 32 / SPLIT;        // LPW rows per warp, SPLIT lanes per row
  extern __shared__ __align__(128) unsigned char tile_smem[];
  // UNI: no per-type table in shared memory (its constants are kernel parameters), which is what lets two CTAs share an SM on
  // the C2 tiles (2 x (3508 slots x 32 B + 0.6 KB) = 226 KB): one evaluates while the other waits at its tile barrier / bulk copies
  constexpr int NK = UNI ? 0 : 1;
  TileSmem<2, NK> S(tile_smem, A.cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!UNI) load_tab(S.T, A.tab[0]);
  if (tid == 0) mbar_init(S.bar, 1);
  const PairTab &T = S.T[0];                                      // (not read when UNI)
  const TileUni &U = A.uni[0];
  const double2 *P0 = S.part, *P1 = S.part + A.cap;
  const int ntiles = A.ntiles, g_far = A.scan_far[0], g_mid = A.scan_far[2];
  TileLoop<0x3, NK, NT> L(A, S, ntiles);
  L.start();
  while (const TileDesc *Dp = L.peek()) {
    const TileDesc &D = *Dp;
    const int zone = A.tzone ? A.tzone[S.tile[L.cur]] : (g_far | (g_mid << 1));
    const int scan_far = zone & 1, scan_mid = zone & 2;
    // (lanes per row chosen by the size of the tile -- 2x / 4x for tiles of <= 128 / <= 64 rows, so that no warp idles at the tile
    // barrier -- was measured in round 2: density 0.246 -> 0.320 ms, force unchanged; the extra lanes per row break the bank-aware
    // entry order, which assumes 8 consecutive rows per quarter-warp.  Fixed split.)
    constexpr int lpw = LPW, split = SPLIT;
    const int sub = lane / lpw, rl = lane % lpw, rpp = (NT / 32) * lpw;
    for (int rb = 0; rb < D.nrows; rb += rpp) {
      const int rt = rb + warp * lpw + rl, row = D.row0 + rt;
      bool valid = rt < D.nrows;
      const int myslot = D.seg_slot[2 * D.center] + (D.row0 - D.seg_src[2 * D.center]) + rt;
      const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
      // global loads of the row first (type, counts, first group), the wait for the tile's records after them
      int ti = 0, nn0 = 0, nf0 = 0; uint4 E0 = make_uint4(0, 0, 0, 0);
      if (valid) {
        ti = tw_type(__double_as_longlong(A.xt[row].w)); nn0 = A.numneigh[row];
        if (scan_far | scan_mid) nf0 = A.numfar[row];
        if (sub < A.ngrp) E0 = ldg_nc_u4(A.near + rbase + sub * 32);
      }
      if (rb == 0) L.wait();
      double2 a = make_double2(0, 0), b = a;
      if (valid) { a = P0[myslot]; b = P1[myslot]; }
      const unsigned rowmask = (unsigned)(U.mapmask >> (ti * 8)) & 0xffu;
      if (valid && (UNI ? rowmask == 0 : T.iskip[ti] != 0)) valid = false;     // atoms of skipped types keep their integrated rho; iskip = no mapped partner type
      double acc = 0.0;
      for (int pass = 0; pass < 3; pass++) {           // near row | mid entries (from the back of the far row) | far row
        if (pass && !(pass == 1 ? scan_mid : scan_far)) continue;
        const int nf = valid ? nf0 : 0;
        const int ng = pass == 0 ? (valid ? (nn0 + 7) >> 3 : 0) : (((pass == 1 ? nf >> 16 : nf & 0xffff) + 7) >> 3);
        const ptrdiff_t dir = pass == 1 ? -32 : 32;
        const uint4 *lp = pass == 0 ? A.near + rbase : (pass == 1 ? A.far + rbase + (size_t)(A.ngrp - 1) * 32 : A.far + rbase);
        uint4 En = E0;
        if (pass && sub < ng) En = ldg_nc_u4(lp + sub * dir);
        for (int gi = sub; gi < ng; gi += split) {
          const uint4 E = En;
          if (gi + split < ng) En = ldg_nc_u4(lp + (gi + split) * dir);     // next group in flight while this one is evaluated
          const unsigned w[4] = {E.x, E.y, E.z, E.w};
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const unsigned ent = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            const int slot = ent & TILE_SLOT_MASK, tj = ent >> TILE_SLOT_BITS;
            const double2 qa = P0[slot], qb = P1[slot];
            const double rsq = rsq_fma(a.x - qa.x, a.y - qa.y, b.x - qb.x);
            if (UNI) {
              const bool hit = (rsq < U.cutsq) & ((rowmask >> tj) & 1u);
              double wf = fma(-rsq, U.c1, 1.0);                   // 1 - r^2/h^2
              wf = wf * wf; wf = wf * wf;
              acc += hit ? wf : 0.0;                              // x mass C_d / h^d at the end
            } else {
              const int ij = ti * MAXT1 + tj;                     // empty entries: cutsq[ti][0] = -1
              if (rsq < T.cutsq[ij]) {
                double wf = 1.0 - rsq * T.c1[ij];
                wf = wf * wf; wf = wf * wf;
                acc += T.mass[tj] * (T.c0[ij] * wf);              // C_d (1-r^2/h^2)^4 / h^d
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = lpw; o < 32; o <<= 1) acc += __shfl_xor_sync(FULLMASK, acc, o);
      if (UNI) acc *= U.mass * U.c0;
      if (valid && sub == 0) A.vr_out[row].w = (UNI ? U.mass * U.self : T.mass[ti] * T.self0[ti]) + acc;
    }
    L.release();
  }
}

// ------------------------------------------------------------ fused forces --
//  K_TAIT   PairSPHTaitwater::compute        pair_sph_taitwater.cpp:101-196
//  K_MORRIS PairSPHTaitwaterMorris::compute  pair_sph_taitwater_morris.cpp:102-196
//  K_HEAT   PairSPHHeatConduction::compute   pair_sph_heatconduction.cpp:76-132
//  K_IDEAL  PairSPHIdealGas::compute         pair_sph_idealgas.cpp:48-175 (taitwater with p/rho^2 = 0.4 e/(m rho), c = sqrt(0.4 e/m))
// UNI: both tables are uniform (TileUni) -> a branch-free body on register constants that ptxas interleaves across the
// 8 neighbors of a group; otherwise the per-type tables are read from shared memory.
// VIR (thermo steps only): every row also sums 1/2 (x_i - x_j) (x) F_ij over its entries.  Summed over all rows this is
// Pair::virial_fdotr_compute's sum of x (x) f over owned + ghost atoms (pair.cpp:1403-1451): a pair of two owned atoms appears in
// both rows, a pair with a ghost contributes its other half on the rank (or periodic image) that owns the ghost.
template <int KINDS, int SPLIT, bool UNI, bool VIR>
__global__ void __launch_bounds__(TILE_ROWS * SPLIT, 1) k_tile_force(const __grid_constant__ TileArgs A)
{
  constexpr bool HAS_FLUID = (KINDS & (K_TAIT | K_MORRIS | K_IDEAL)) != 0;
  constexpr bool HAS_HEAT = (KINDS & K_HEAT) != 0;
  constexpr int NK = (HAS_FLUID ? 1 : 0) + (HAS_HEAT ? 1 : 0);
  constexpr int NPARTS = HAS_FLUID ? (HAS_HEAT ? 5 : 4) : 3;
  constexpr int PE = HAS_FLUID ? 4 : 2;                           // part holding e
  constexpr int I_HEAT = HAS_FLUID ? 1 : 0;
  constexpr int NT = TILE_ROWS * SPLIT, LPW = 32 / SPLIT;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  TileSmem<NPARTS, NK> S(tile_smem, A.cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = 0; t < NK; t++) load_tab(S.T + t, A.tab[t]);
  if (tid == 0) mbar_init(S.bar, 1);
  const double2 *P0 = S.part, *P1 = P0 + A.cap, *P2 = P1 + A.cap, *P3 = P2 + A.cap, *PEp = S.part + (size_t)PE * A.cap;
  const TileUni &UF = A.uni[0], &UH = A.uni[I_HEAT];
  // register constants of the uniform body
  const double u_eta = 0.01 * UF.h * UF.h, u_vch = -UF.visc * (UF.cs + UF.cs) * UF.h, u_vci = -UF.visc * UF.h;
  const double u_k1 = -UF.mass * UF.mass * UF.c0, u_k2 = 2.0 * UF.visc * UF.mass * UF.mass * UF.c0, u_k3 = UF.mass * UF.c0;
  const double u_heat = HAS_HEAT ? 2.0 * UH.mass * UH.mass * UH.visc / (UH.mass + UH.mass) * UH.c0 : 0.0;
  const int ntiles = A.ntiles, g_far = A.scan_far[0], g_mid = A.scan_far[2];
  TileLoop<(1 << NPARTS) - 1, NK, NT> L(A, S, ntiles);
  L.start();
  while (const TileDesc *Dp = L.peek()) {
    const TileDesc &D = *Dp;
    const int zone = A.tzone ? A.tzone[S.tile[L.cur]] : (g_far | (g_mid << 1));
    const int scan_far = zone & 1, scan_mid = zone & 2;
    constexpr int lpw = LPW, split = SPLIT;      // see k_tile_rhosum
    const int sub = lane / lpw, rl = lane % lpw, rpp = (NT / 32) * lpw;
    for (int rb = 0; rb < D.nrows; rb += rpp) {
      const int rt = rb + warp * lpw + rl, row = D.row0 + rt;
      const bool valid = rt < D.nrows;
      const int myslot = D.seg_slot[2 * D.center] + (D.row0 - D.seg_src[2 * D.center]) + rt;
      const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
      // global loads of the row first (type, counts, first group), the wait for the tile's records after them
      int ti = 0, nn0 = 0, nf0 = 0; uint4 E0 = make_uint4(0, 0, 0, 0);
      if (valid) {
        ti = tw_type(__double_as_longlong(A.xt[row].w)); nn0 = A.numneigh[row];
        if (scan_far | scan_mid) nf0 = A.numfar[row];
        if (sub < A.ngrp) E0 = ldg_nc_u4(A.near + rbase + sub * 32);
      }
      if (rb == 0) L.wait();
      double2 a = make_double2(0, 0), b = make_double2(0, 1), c = a, d = a; double ei = 0.0;
      if (valid) {
        a = P0[myslot]; b = P1[myslot];
        if (HAS_FLUID) { c = P2[myslot]; d = P3[myslot]; }
        if (HAS_HEAT) ei = PEp[myslot].x;
      }
      const double rhoi = b.y, mi = S.T[0].mass[ti];
      const double ci = (KINDS & K_IDEAL) ? sqrt(fmax(d.y * rhoi, 0.0)) : 0.0;
      const unsigned maskf = (unsigned)(UF.mapmask >> (ti * 8)) & 0xffu, maskh = (unsigned)(UH.mapmask >> (ti * 8)) & 0xffu;
      double fx = 0, fy = 0, fz = 0, adrho = 0, ade = 0;
      double u_drho = 0, u_de = 0, u_deh = 0;                                 // uniform body: raw sums, scaled after the loop
      double w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0;                   // VIR: xx yy zz xy xz yz
      auto vir = [&](double dx, double dy, double dz, double Fx, double Fy, double Fz) {
        w0 += dx * Fx; w1 += dy * Fy; w2 += dz * Fz; w3 += dx * Fy; w4 += dx * Fz; w5 += dy * Fz;
      };
      for (int pass = 0; pass < 3; pass++) {           // near row | mid entries (from the back of the far row) | far row
        if (pass && !(pass == 1 ? scan_mid : scan_far)) continue;
        const int nf = valid ? nf0 : 0;
        const int ng = pass == 0 ? (valid ? (nn0 + 7) >> 3 : 0) : (((pass == 1 ? nf >> 16 : nf & 0xffff) + 7) >> 3);
        const ptrdiff_t dir = pass == 1 ? -32 : 32;
        const uint4 *lp = pass == 0 ? A.near + rbase : (pass == 1 ? A.far + rbase + (size_t)(A.ngrp - 1) * 32 : A.far + rbase);
        uint4 En = E0;
        if (pass && sub < ng) En = ldg_nc_u4(lp + sub * dir);
        for (int gi = sub; gi < ng; gi += split) {
          const uint4 E = En;
          if (gi + split < ng) En = ldg_nc_u4(lp + (gi + split) * dir);     // next group in flight while this one is evaluated
          const unsigned w[4] = {E.x, E.y, E.z, E.w};
#pragma unroll kForceUnroll
          for (int e = 0; e < 8; e++) {
            const unsigned ent = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            const int slot = ent & TILE_SLOT_MASK, tj = ent >> TILE_SLOT_BITS;
            const double2 qa = P0[slot], qb = P1[slot];
            const double dx = a.x - qa.x, dy = a.y - qa.y, dz = b.x - qb.x;
            const double rhoj = qb.y;
            if (UNI) {
              const double rsq = rsq_fma(dx, dy, dz);
              // r = sqrt(rsq) to ~1e-13 (fast_sqrt13).  The MUFU seed is clamped to a finite value and halved on the integer
              // pipe, so coincident particles give r = 0 (not 0 * inf) and flow through the formulas as in the reference,
              // which has no division by r either
              double y;
              asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(rsq));
              const int yhi = min(__double2hiint(y), 0x7fe00000);
              y = __hiloint2double(yhi, 0);                                   // (the seed's low word is zero)
              const double g = rsq * y;
              const double r = fma(g, fma(-__hiloint2double(yhi - 0x00100000, 0), g, 0.5), g);
              if (HAS_FLUID) {
                // inside the cutoff <=> h - r >= 0, read off the sign bit (integer pipe; cut = h for these styles and the
                // weight (h - r)^2 vanishes there, so a pair within rounding of the cutoff adds ~1e-26 h^2 either way)
                const double hr = UF.h - r;
                const int hit = (~__double2hiint(hr) >> 31) & (int)(maskf >> tj) & 1;
                const double2 qc = P2[slot], qd = P3[slot];
                double wfd = hr * hr;                                         // Lucy (dW/dr)/r = c0 (h - r)^2  (:135-151), c0 folded into the constants
                wfd = hit ? wfd : 0.0;
                const double dvx = c.x - qc.x, dvy = c.y - qc.y, dvz = d.x - qd.x;
                const double dvdr = dx * dvx + dy * dvy + dz * dvz;
                if (KINDS & (K_TAIT | K_IDEAL)) {
                  // sph/idealgas: the sound speeds are per particle, c = sqrt(0.4 e / m) = sqrt((p/rho^2) rho)
                  const double vch = (KINDS & K_IDEAL) ? u_vci * (ci + fast_sqrt(fmax(qd.y * rhoj, 0.0))) : u_vch;
                  // Monaghan artificial viscosity (:163-169), only for approaching pairs (sign bit of dvdr, integer pipe)
                  double nv = vch * dvdr;
                  nv = __double2hiint(dvdr) < 0 ? nv : 0.0;
                  const double den = (rsq + u_eta) * (rhoi + rhoj);
                  double yr;
                  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yr) : "d"(den));
                  yr = fma(yr, fma(-den, yr, 1.0), yr);                       // one Newton step on the MUFU seed: ~1e-13, as the sqrt above
                  const double fpair = fma(nv, yr, d.y + qd.y) * wfd;         // x -m m c0 after the loop
                  fx = fma(dx, fpair, fx); fy = fma(dy, fpair, fy); fz = fma(dz, fpair, fz);
                  if (VIR) vir(dx, dy, dz, dx * fpair, dy * fpair, dz * fpair);
                  u_de = fma(fpair, dvdr, u_de);                              // x -0.5 (-m m c0) at the end
                } else {                                                      // Morris viscosity (morris :165-176)
                  const double fvisc = fast_div(u_k2 * wfd, rhoi * rhoj);     // 2 mu m m c0 (h - r)^2 / (rho_i rho_j)
                  const double fpair = u_k1 * (d.y + qd.y) * wfd;
                  fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
                  if (VIR) vir(dx, dy, dz, dx * fpair + dvx * fvisc, dy * fpair + dvy * fvisc, dz * fpair + dvz * fvisc);
                  u_de += fpair * dvdr + fvisc * (dvx * dvx + dvy * dvy + dvz * dvz);
                }
                u_drho = fma(dvdr, wfd, u_drho);                              // x m c0 at the end
              }
              if (HAS_HEAT) {
                double wfd = UH.h - r;
                const int hit = (~__double2hiint(wfd) >> 31) & (int)(maskh >> tj) & 1;
                wfd = wfd * wfd;
                wfd = hit ? wfd : 0.0;
                const double ej = PEp[slot].x;
                // 2 mi mj/(mi+mj) (rho_i+rho_j)/(rho_i rho_j) D (e_i - e_j) W'/r  (:122-125), one division
                u_deh += fast_div((rhoi + rhoj) * (ei - ej) * wfd, rhoi * rhoj);
              }
            } else {
              const double rsq = rsq_fma(dx, dy, dz);
              const int ij = ti * MAXT1 + tj;                                 // empty entries: cutsq[ti][0] = -1
              bool any = false;
#pragma unroll
              for (int t = 0; t < NK; t++) any |= rsq < S.T[t].cutsq[ij];
              if (!any) continue;
              const double mj = S.T[0].mass[tj];
              const double rinv = rsqrt(rsq), r = rsq > 0.0 ? rsq * rinv : 0.0;      // coincident particles: r = 0 as the reference's sqrt gives (not 0 * inf)
              if (HAS_FLUID) {
                const PairTab &P = S.T[0];
                if (rsq < P.cutsq[ij]) {
                  const double2 qc = P2[slot], qd = P3[slot];
                  const double h = P.h[ij];
                  double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;
                  const double dvx = c.x - qc.x, dvy = c.y - qc.y, dvz = d.x - qd.x;
                  const double dvdr = dx * dvx + dy * dvy + dz * dvz;
                  const double mm = mi * mj;
                  if (KINDS & (K_TAIT | K_IDEAL)) {
                    double fvisc = 0.0;
                    if (dvdr < 0.0) {
                      const double cc = (KINDS & K_IDEAL) ? ci + sqrt(qd.y * rhoj) : P.cs[ti] + P.cs[tj];
                      fvisc = -P.visc[ij] * cc * (h * dvdr) / ((rsq + 0.01 * h * h) * (rhoi + rhoj));
                    }
                    const double fpair = -mm * (d.y + qd.y + fvisc) * wfd;
                    fx += dx * fpair; fy += dy * fpair; fz += dz * fpair;
                    if (VIR) vir(dx, dy, dz, dx * fpair, dy * fpair, dz * fpair);
                    ade += -0.5 * fpair * dvdr;
                  } else {
                    const double fvisc = 2.0 * P.visc[ij] / (rhoi * rhoj) * mm * wfd;
                    const double fpair = -mm * (d.y + qd.y) * wfd;
                    fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
                    if (VIR) vir(dx, dy, dz, dx * fpair + dvx * fvisc, dy * fpair + dvy * fvisc, dz * fpair + dvz * fvisc);
                    ade += -0.5 * (fpair * dvdr + fvisc * (dvx * dvx + dvy * dvy + dvz * dvz));
                  }
                  adrho += mj * dvdr * wfd;
                }
              }
              if (HAS_HEAT) {
                const PairTab &P = S.T[I_HEAT];
                if (rsq < P.cutsq[ij]) {
                  const double h = P.h[ij];
                  double wfd = h - r; wfd = P.c0[ij] * wfd * wfd;
                  const double ej = PEp[slot].x;
                  ade += 2.0 * mi * mj * (rhoi + rhoj) * P.visc[ij] * (ei - ej) * wfd / ((mi + mj) * (rhoi * rhoj));
                }
              }
            }
          }
        }
      }
      if (UNI) {
        if (KINDS & (K_TAIT | K_IDEAL)) {                                       // the pair sums above are without the factor -m m c0
          fx *= u_k1; fy *= u_k1; fz *= u_k1; u_de *= u_k1;
          if (VIR) { w0 *= u_k1; w1 *= u_k1; w2 *= u_k1; w3 *= u_k1; w4 *= u_k1; w5 *= u_k1; }
        }
        adrho = u_k3 * u_drho; ade = -0.5 * u_de + u_heat * u_deh;
      }
#pragma unroll
      for (int o = lpw; o < 32; o <<= 1) {
        fx += __shfl_xor_sync(FULLMASK, fx, o); fy += __shfl_xor_sync(FULLMASK, fy, o); fz += __shfl_xor_sync(FULLMASK, fz, o);
        adrho += __shfl_xor_sync(FULLMASK, adrho, o); ade += __shfl_xor_sync(FULLMASK, ade, o);
        if (VIR) {
          w0 += __shfl_xor_sync(FULLMASK, w0, o); w1 += __shfl_xor_sync(FULLMASK, w1, o); w2 += __shfl_xor_sync(FULLMASK, w2, o);
          w3 += __shfl_xor_sync(FULLMASK, w3, o); w4 += __shfl_xor_sync(FULLMASK, w4, o); w5 += __shfl_xor_sync(FULLMASK, w5, o);
        }
      }
      if (valid && sub == 0) {
        double4 f = make_double4(0, 0, 0, 0); double de0 = 0.0;
        if (A.accum) { f = A.fd[row]; de0 = A.de[row]; }          // first force pass since force_clear: both are zero, no read-modify-write
        f.x += fx; f.y += fy; f.z += fz; f.w += adrho;
        A.fd[row] = f;
        A.de[row] = de0 + ade;
        if (VIR) {
          double *w = A.virow + (size_t)row * 6;
          w[0] += 0.5 * w0; w[1] += 0.5 * w1; w[2] += 0.5 * w2; w[3] += 0.5 * w3; w[4] += 0.5 * w4; w[5] += 0.5 * w5;
        }
      }
    }
    L.release();
  }
}

// ======================================================================= multiphase styles on tiles ====
// Record parts (tile order, ghosts keep their possibly one-step-stale rho / colorgradient):
//   P0 x,y   P1 z,rho   P2 vest.x,vest.y   P3 vest.z, pressure   P4 V^2 = (m/rho)^2, T = e/cv   P7 m,0
//   P5 Axx,Ayy   P6 Axy,Azz   P8 Axz,Ayz (3-D)      A = V^2 Pi, the particle's surface stress
// colorgradient pass: P0 x,y   P1 z, V^2
// PairSPHSurfaceTension::compute (pair_sph_surfacetension.cpp:140-169) spells out, pair by pair, the product of the unit vector e_ij
// with Pi = (|c|^2 / d  I - c (x) c) / |c|  (c = colorgradient, d = dimension, Pi = 0 where |c| <= EPSILON) of either particle and adds
// (Pi_i e V_i^2 + Pi_j e V_j^2) dW/dr.  Pi depends on one particle only, so it is formed once per particle here and the pair loop is
// left with (A_i + A_j) e: 19 fp64 instructions per pair instead of 72 (profiles/r02_mpforce_*: the pass is bound by the fp64 pipe).
struct TileRecMpArgs {
  int nlocal, nall, pstride, mode, dim;  // mode 0: colorgradient records, 1: force records
  const int *gorder; const double4 *xt, *vr, *cgm; const double *e, *cv; const PairTab *fluid;
  double2 *rec;
};
__global__ void k_tile_records_mp(TileRecMpArgs A)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.nall) return;
  int src = i < A.nlocal ? i : A.nlocal + A.gorder[i - A.nlocal];
  double4 x = A.xt[src], v = A.vr[src], c = A.cgm[src];
  const double rho = v.w, V = c.w / rho;
  const size_t ps = A.pstride;
  A.rec[i] = make_double2(x.x, x.y);
  if (A.mode == 0) { A.rec[ps + i] = make_double2(x.z, V * V); return; }
  int t = tw_type(__double_as_longlong(x.w));
  // pair_sph_taitwater_multiphase.cpp:289-292
  double P = A.fluid ? A.fluid->B[t] * (pow(rho / A.fluid->rho0[t], A.fluid->gamma[t]) - A.fluid->rb[t]) : 0.0;
  A.rec[ps + i] = make_double2(x.z, rho);
  A.rec[2 * ps + i] = make_double2(v.x, v.y);
  A.rec[3 * ps + i] = make_double2(v.z, P);
  A.rec[4 * ps + i] = make_double2(V * V, A.e[src] / A.cv[src]);
  A.rec[7 * ps + i] = make_double2(c.w, 0.0);
  const double cxx = c.x * c.x, cyy = c.y * c.y, czz = c.z * c.z;
  if (A.dim == 3) {                      // (:153-169)
    const double o3 = 0.3333333333333333, t3 = 0.6666666666666666;
    const double a = sqrt(cxx + cyy + czz), s = (a > EPSILON_CG ? 1.0 / a : 0.0) * (V * V);
    A.rec[5 * ps + i] = make_double2((o3 * czz + o3 * cyy - t3 * cxx) * s, (o3 * czz - t3 * cyy + o3 * cxx) * s);
    A.rec[6 * ps + i] = make_double2(-(c.x * c.y) * s, (-t3 * czz + o3 * cyy + o3 * cxx) * s);
    A.rec[8 * ps + i] = make_double2(-(c.x * c.z) * s, -(c.y * c.z) * s);
  } else {                               // (:140-151): the 2-D norm and trace
    const double a = sqrt(cxx + cyy), s = (a > EPSILON_CG ? 1.0 / a : 0.0) * (V * V), h2 = (cyy + cxx) / 2;
    A.rec[5 * ps + i] = make_double2((h2 - cxx) * s, (h2 - cyy) * s);
    A.rec[6 * ps + i] = make_double2(-(c.x * c.y) * s, 0.0);
  }
}

// branch-free quintic spline (sph_kernel_quintic.cpp:17-73, argument q = 3 r / h, without the norm): the clamped form
// max(3-q,0)^5 - 6 max(2-q,0)^5 + 15 max(1-q,0)^5 is the same piecewise polynomial as quintic_w / quintic_dw
// max(x, 0) on the integer pipe (the fp64 pipe is what binds these kernels): the sign bit masks both words
__device__ __forceinline__ double dclamp0(double x)
{
  const int hi = __double2hiint(x), m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double quintic_w_bf(double q)
{
  const double a = dclamp0(3.0 - q), b = dclamp0(2.0 - q), c = dclamp0(1.0 - q);
  const double a2 = a * a, b2 = b * b, c2 = c * c;
  return fma(15.0 * c, c2 * c2, fma(-6.0 * b, b2 * b2, a * (a2 * a2)));
}
// dW/dq / (-5) = max(3-q,0)^4 - 6 max(2-q,0)^4 + 15 max(1-q,0)^4: the callers fold the factor -5 into their constants
__device__ __forceinline__ double quintic_dw5_bf(double q)
{
  const double a = dclamp0(3.0 - q), b = dclamp0(2.0 - q), c = dclamp0(1.0 - q);
  const double a2 = a * a, b2 = b * b, c2 = c * c;
  return fma(15.0 * c2, c2, fma(-6.0 * b2, b2, a2 * a2));
}
__device__ __forceinline__ double quintic_dw_bf(double q) { return -5.0 * quintic_dw5_bf(q); }
// r = sqrt(a) and 1/r, branch-free (see fast_sqrt); NaN / inf for a = 0, which the callers mask with dpos()
__device__ __forceinline__ void fast_sqrt_rinv(double a, double &r, double &rinv)
{ // one coupled Newton step from the 2^-22 seed: both results to ~1e-13 (see fast_sqrt13); the seed is halved on the integer pipe
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = a * y, hh = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  const double e = fma(-hh, g, 0.5);
  r = fma(g, e, g); rinv = fma(y, e, y);
}

// rows per pass and lanes per row of a 512-thread stage CTA: 256 rows x 2 lanes, or 128 rows x 4 lanes for small tiles
#define TILE_MP_NT (TILE_ROWS * 2)

// MPK = 0: PairSPHRhoSumMultiphase::compute  pair_sph_rhosum_multiphase.cpp:113-168 (quintic, number density * own mass)
// MPK = 1: PairSPHColorGradient::compute    pair_sph_colorgradient.cpp:119-184
// Full lists: every entry of an owned row counts, whatever its ownership flags.
// GU: cutoff and kernel constants are the same for every mapped type pair (TileUni) -> registers; else the per-pair tables.
template <int MPK, bool GU>
// (256-thread CTAs, three per SM, measured against 512-thread CTAs, two per SM, on the C3 styles: density 0.426 -> 0.415 ms, colorgradient 0.502 -> 0.483 ms)
#ifndef TILE_MPFULL_NT
#define TILE_MPFULL_NT 256
#endif
__global__ void __launch_bounds__(TILE_MPFULL_NT, TILE_MPFULL_NT <= 256 ? 3 : 2) k_tile_full_mp(const __grid_constant__ TileArgs A)
{
  constexpr int NT = TILE_MPFULL_NT;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  TileSmem<2, 1> S(tile_smem, A.cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_tab(S.T, A.tab[0]);
  if (tid == 0) mbar_init(S.bar, 1);
  const PairTab &T = S.T[0];
  const TileUni &U = A.uni[0];
  const double2 *P0 = S.part, *P1 = S.part + A.cap;
  const int ntiles = A.ntiles, scan_far = A.scan_far[0], scan_mid = A.scan_far[2];
  TileLoop<0x3, 1, NT> L(A, S, ntiles);
  L.start();
  while (const TileDesc *Dp = L.peek()) {
    const TileDesc &D = *Dp;
    const int lpw = D.nrows <= TILE_ROWS / 2 ? 8 : 16, split = 32 / lpw, rpp = (NT / 32) * lpw;
    const int sub = lane / lpw, rl = lane % lpw;
    for (int rb = 0; rb < D.nrows; rb += rpp) {
      const int rt = rb + warp * lpw + rl, row = D.row0 + rt;
      bool valid = rt < D.nrows;
      const int myslot = D.seg_slot[2 * D.center] + (D.row0 - D.seg_src[2 * D.center]) + rt;
      const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
      // global loads of the row first (type, counts, first group), the wait for the tile's records after them
      int ti = 0, nn0 = 0, nf0 = 0; uint4 E0 = make_uint4(0, 0, 0, 0);
      if (valid) {
        ti = tw_type(__double_as_longlong(A.xt[row].w)); nn0 = A.numneigh[row];
        if (scan_far | scan_mid) nf0 = A.numfar[row];
        if (sub < A.ngrp) E0 = ldg_nc_u4(A.near + rbase + sub * 32);
      }
      if (rb == 0) L.wait();
      double2 a = make_double2(0, 0), b = a;
      if (valid) { a = P0[myslot]; b = P1[myslot]; }
      if (valid && T.iskip[ti]) valid = false;
      const unsigned rowmask = (unsigned)(U.mapmask >> (ti * 8)) & 0xffu;
      double acc = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
      for (int pass = 0; pass < 3; pass++) {           // near row | mid entries (from the back of the far row) | far row
        if (pass && !(pass == 1 ? scan_mid : scan_far)) continue;
        const int nf = valid ? nf0 : 0;
        const int ng = pass == 0 ? (valid ? (nn0 + 7) >> 3 : 0) : (((pass == 1 ? nf >> 16 : nf & 0xffff) + 7) >> 3);
        const ptrdiff_t dir = pass == 1 ? -32 : 32;
        const uint4 *lp = pass == 0 ? A.near + rbase : (pass == 1 ? A.far + rbase + (size_t)(A.ngrp - 1) * 32 : A.far + rbase);
        uint4 En = E0;
        if (pass && sub < ng) En = ldg_nc_u4(lp + sub * dir);
        for (int gi = sub; gi < ng; gi += split) {
          const uint4 E = En;
          if (gi + split < ng) En = ldg_nc_u4(lp + (gi + split) * dir);
          const unsigned w[4] = {E.x, E.y, E.z, E.w};
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const unsigned ent = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            const int slot = ent & TMP_SLOT_MASK, tj = ent >> TILE_SLOT_BITS, ij = ti * MAXT1 + tj;
            const double2 qa = P0[slot], qb = P1[slot];
            const double dx = a.x - qa.x, dy = a.y - qa.y, dz = b.x - qb.x;
            const double rsq = rsq_fma(dx, dy, dz);
            const bool hit = GU ? ((rsq < U.cutsq) & ((rowmask >> tj) & 1u) & dpos(rsq)) : ((rsq < T.cutsq[ij]) & dpos(rsq));   // empty entries: type 0 is never mapped
            const double c0 = GU ? U.c0 : T.c0[ij], c1 = GU ? U.c1 : T.c1[ij];
            if (MPK == 0) {
              const double wv = quintic_w_bf(3.0 * (fast_sqrt13(rsq) * c1));
              acc += hit ? (GU ? wv : c0 * wv) : 0.0;
            } else {
              double r, rinv; fast_sqrt_rinv(rsq, r, rinv);
              const double wfd = quintic_dw_bf(3.0 * (r * c1)) * c0;                 // dW/dr
              double sc = -wfd * T.visc[ij] * qb.y * rinv;                           // -W' alpha / sigma_j^2 / r   (sigma_i applied at the end)
              sc = hit ? sc : 0.0;
              ax += sc * dx; ay += sc * dy; az += sc * dz;
            }
          }
        }
      }
      for (int o = lpw; o < 32; o <<= 1) {
        if (MPK == 0) acc += __shfl_xor_sync(FULLMASK, acc, o);
        else { ax += __shfl_xor_sync(FULLMASK, ax, o); ay += __shfl_xor_sync(FULLMASK, ay, o); az += __shfl_xor_sync(FULLMASK, az, o); }
      }
      if (valid && sub == 0) {
        if (MPK == 0) A.vr_out[row].w = (T.self0[ti] + (GU ? U.c0 * acc : acc)) * A.vm[row].w;   // rho[i] *= imass (:170)
        else {
          const double sigmai = A.vr_out[row].w / A.vm[row].w;
          double4 c = A.cg_out[row];
          c.x = ax * sigmai; c.y = ay * sigmai; c.z = (A.dim == 3) ? az * sigmai : 0.0;
          A.cg_out[row] = c;
        }
      }
    }
    L.release();
  }
}

//  K_TAITMP PairSPHTaitwaterMultiphase::compute  pair_sph_taitwater_multiphase.cpp:103-182
//  K_SURF   PairSPHSurfaceTension::compute       pair_sph_surfacetension.cpp:81-190
//  K_HEATMP PairSPHHeatConductionMultiPhase      pair_sph_heatconduction_multiphase.cpp:78-127
//  K_HEATPC PairSPHHeatConductionPhaseChange     pair_sph_heatconduction_phasechange.cpp:83-139
// Tiles of owned rows and tiles of ghost rows (D.ghost) run through the same kernel: a ghost row accumulates what the
// reference adds to the ghost atom (reverse-communicated afterwards); an owned row skips the pairs with a ghost that the
// other side owns.  See b200_pair.cuh for the two orientation quirks that use the ownership bit.
// The body is branch-free (predicated contributions) so that ptxas interleaves the neighbors of a group.
// GU: every sub-style has one cutoff / one set of kernel constants for all its mapped type pairs AND all sub-styles share
// the smoothing length (the usual deck) -> one kernel-derivative evaluation per pair on register constants.  The cutoff of
// these styles is the support of the spline (cut = h, fill_tab), whose clamped form is exactly 0 beyond it, so the GU body
// needs no cutoff compare: an entry outside contributes 0.0 by itself.
template <int KINDS, bool DIM3> struct MpParts {
  static constexpr int mask = 0x03 | ((KINDS & K_TAITMP) ? 0x1c : 0) | ((KINDS & K_SURF) ? (DIM3 ? 0x160 : 0x60) : 0) | ((KINDS & (K_HEATMP | K_HEATPC)) ? 0x90 : 0);
  static constexpr int n = __builtin_popcount(mask);
  static constexpr int nk = ((KINDS & K_TAITMP) ? 1 : 0) + ((KINDS & K_SURF) ? 1 : 0) + ((KINDS & (K_HEATMP | K_HEATPC)) ? 1 : 0);
  __host__ __device__ static constexpr int idx(int p) { return __builtin_popcount(mask & ((1 << p) - 1)); }
};
template <int KINDS, bool DIM3, bool GU>
__global__ void __launch_bounds__(TILE_MP_NT, 1) k_tile_force_mp(const __grid_constant__ TileArgs A)
{
  using MP = MpParts<KINDS, DIM3>;
  constexpr bool HAS_FLUID = (KINDS & K_TAITMP) != 0, HAS_SURF = (KINDS & K_SURF) != 0, HAS_HEAT = (KINDS & (K_HEATMP | K_HEATPC)) != 0;
  constexpr int NK = MP::nk, NP = MP::n;
  constexpr int I_FLUID = 0, I_SURF = HAS_FLUID ? 1 : 0, I_HEAT = I_SURF + (HAS_SURF ? 1 : 0);
  constexpr bool WRITES_DE = HAS_HEAT;
  constexpr int NT = TILE_MP_NT;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  TileSmem<NP, NK> S(tile_smem, A.cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = 0; t < NK; t++) load_tab(S.T + t, A.tab[t]);
  if (tid == 0) mbar_init(S.bar, 1);
  const PairTab *T = S.T;
  auto part = [&](int p) { return S.part + (size_t)MP::idx(p) * A.cap; };
  const double2 *P0 = part(0), *P1 = part(1), *P2 = part(2), *P3 = part(3), *P4 = part(4), *P5 = part(5), *P6 = part(6), *P7 = part(7), *P8 = part(8);
  const int ntiles = A.ntiles, scan_far = A.scan_far[0], scan_mid = A.scan_far[2];
  const size_t ps = A.pstride;
  // GU: q = 3 r / h and the factor -5 of the spline derivative folded into register constants
  const double gu_q = 3.0 * A.uni[0].c1;
  double gu_c0[3] = {0, 0, 0};
#pragma unroll
  for (int t = 0; t < NK; t++) gu_c0[t] = -5.0 * A.uni[t].c0;
  TileLoop<MP::mask, NK, NT> L(A, S, ntiles);
  L.start();
  while (const TileDesc *Dp = L.peek()) {
    const TileDesc &D = *Dp;
    const bool ghostrow = D.ghost != 0;
    const int lpw = D.nrows <= TILE_ROWS / 2 ? 8 : 16, split = 32 / lpw, rpp = (NT / 32) * lpw;
    const int sub = lane / lpw, rl = lane % lpw;
    for (int rb = 0; rb < D.nrows; rb += rpp) {
      const int rt = rb + warp * lpw + rl, row = D.row0 + rt;                  // tile-order index
      const bool valid = rt < D.nrows;
      const int rr = valid ? row : D.row0;
      // the row particle's own record (global, coalesced): ghost rows are not among the staged candidates
      const double2 a = A.rec[rr], b = A.rec[ps + rr];
      double2 c = make_double2(0, 0), d = c, v4 = c, a5 = c, a6 = c, a8 = c;
      if (HAS_FLUID) { c = A.rec[2 * ps + rr]; d = A.rec[3 * ps + rr]; }
      if (HAS_FLUID || HAS_HEAT) v4 = A.rec[4 * ps + rr];
      if (HAS_SURF) { a5 = A.rec[5 * ps + rr]; a6 = A.rec[6 * ps + rr]; if (DIM3) a8 = A.rec[8 * ps + rr]; }
      const int dev = ghostrow ? A.nlocal + A.gorder[rr - A.nlocal] : rr;     // device index
      const int ti = valid ? tw_type(__double_as_longlong(A.xt[dev].w)) : 0;
      const size_t rbase = (size_t)(row >> 5) * A.ngrp * 32 + (row & 31);
      int nn0 = 0, nf0 = 0; uint4 E0 = make_uint4(0, 0, 0, 0);
      if (valid) {
        nn0 = A.numneigh[row];
        if (scan_far | scan_mid) nf0 = A.numfar[row];
        if (sub < A.ngrp) E0 = ldg_nc_u4(A.near + rbase + sub * 32);
      }
      if (rb == 0) L.wait();                  // everything above came from global memory: its latency overlaps the tile's bulk copies
      const double rhoi = b.y;
      unsigned rmask[3] = {0, 0, 0};
#pragma unroll
      for (int t = 0; t < NK; t++) rmask[t] = (unsigned)(A.uni[t].mapmask >> (ti * 8)) & 0xffu;
      double fx = 0, fy = 0, fz = 0, ade = 0;
      for (int pass = 0; pass < 3; pass++) {           // near row | mid entries (from the back of the far row) | far row
        if (pass && !(pass == 1 ? scan_mid : scan_far)) continue;
        const int nf = valid ? nf0 : 0;
        const int ng = pass == 0 ? (valid ? (nn0 + 7) >> 3 : 0) : (((pass == 1 ? nf >> 16 : nf & 0xffff) + 7) >> 3);
        const ptrdiff_t dir = pass == 1 ? -32 : 32;
        const uint4 *lp = pass == 0 ? A.near + rbase : (pass == 1 ? A.far + rbase + (size_t)(A.ngrp - 1) * 32 : A.far + rbase);
        uint4 En = E0;
        if (pass && sub < ng) En = ldg_nc_u4(lp + sub * dir);
        for (int gi = sub; gi < ng; gi += split) {
          const uint4 E = En;
          if (gi + split < ng) En = ldg_nc_u4(lp + (gi + split) * dir);
          const unsigned w[4] = {E.x, E.y, E.z, E.w};
#pragma unroll kMpForceUnroll
          for (int e = 0; e < 8; e++) {
            const unsigned ent = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            // an owned row skips the ghost pairs the other side owns; empty entries have type 0 (mapped nowhere, cutsq = -1)
            const bool live = ghostrow | ((ent & (TMP_GHOST | TMP_OWNER)) != TMP_GHOST);
            const bool row_owns = !ghostrow & ((ent & TMP_OWNER) != 0);
            const int slot = ent & TMP_SLOT_MASK, tj = ent >> TILE_SLOT_BITS, ij = ti * MAXT1 + tj;
            const double2 qa = P0[slot], qb = P1[slot];
            double2 q4 = make_double2(0, 0);
            if (HAS_FLUID || HAS_HEAT) q4 = P4[slot];
            const double dx = a.x - qa.x, dy = a.y - qa.y, dz = b.x - qb.x;
            const double rsq = rsq_fma(dx, dy, dz);
            const bool ok = live & dpos(rsq);
            const double rhoj = qb.y;
            double r, rinv; fast_sqrt_rinv(rsq, r, rinv);
            double dwq = 0.0;                                                  // (dW/dq) / r without the norm, shared by the sub-styles when GU
            if (GU) dwq = quintic_dw5_bf(r * gu_q) * rinv;
            if (HAS_FLUID) {
              const PairTab &P = T[I_FLUID];
              const bool hit = ok & (GU ? ((rmask[I_FLUID] >> tj) & 1u) != 0 : (rsq < P.cutsq[ij]));
              const double2 qc = P2[slot], qd = P3[slot];
              double wfd = GU ? dwq * gu_c0[I_FLUID] : quintic_dw_bf(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;   // (dW/dr)/r (:137-143)
              wfd = hit ? wfd : 0.0;
              double Pi = d.y, Pj = qd.y;
              if (!P.gamma_uniform) {                                      // p_j uses gamma of the list owner (:148)
                const int to = row_owns ? ti : tj;
                Pi = P.B[ti] * (pow(rhoi / P.rho0[ti], P.gamma[to]) - P.rb[ti]);
                Pj = P.B[tj] * (pow(rhoj / P.rho0[tj], P.gamma[to]) - P.rb[tj]);
              }
              const double pij = fast_div15(rhoj * Pi + rhoi * Pj, rhoi + rhoj);
              const double vw = (v4.x + q4.x) * wfd;
              const double fvisc = vw * P.visc[ij], fpair = -(vw * pij);
              const double dvx = c.x - qc.x, dvy = c.y - qc.y, dvz = d.x - qd.x;
              fx += dx * fpair + dvx * fvisc; fy += dy * fpair + dvy * fvisc; fz += dz * fpair + dvz * fvisc;
            }
            if (HAS_SURF) {
              const PairTab &P = T[I_SURF];
              const bool hit = ok & (GU ? ((rmask[I_SURF] >> tj) & 1u) != 0 : (rsq < P.cutsq[ij]));
              const double2 q5 = P5[slot], q6 = P6[slot];
              double ws = GU ? dwq * gu_c0[I_SURF] : quintic_dw_bf(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;    // (dW/dr)/r (:117-123); e = d / r
              ws = hit ? ws : 0.0;                                         // rinv is inf for a coincident pair: keep 0 * inf out of the sums
              // (A_i + A_j) d, A symmetric
              const double mxx = a5.x + q5.x, myy = a5.y + q5.y, mxy = a6.x + q6.x;
              if (DIM3) {
                const double2 q8 = P8[slot];
                const double mzz = a6.y + q6.y, mxz = a8.x + q8.x, myz = a8.y + q8.y;
                fx += (mxx * dx + mxy * dy + mxz * dz) * ws;
                fy += (mxy * dx + myy * dy + myz * dz) * ws;
                fz += (mxz * dx + myz * dy + mzz * dz) * ws;
              } else {
                fx += (mxx * dx + mxy * dy) * ws;
                fy += (mxy * dx + myy * dy) * ws;
              }
            }
            if (HAS_HEAT) {
              const PairTab &P = T[I_HEAT];
              const bool hit = ok & (GU ? ((rmask[I_HEAT] >> tj) & 1u) != 0 : (rsq < P.cutsq[ij]));
              const double mj = P7[slot].x;
              double wfd = GU ? dwq * gu_c0[I_HEAT] : quintic_dw_bf(3.0 * (r * P.c1[ij])) * P.c0[ij] * rinv;
              double Ti = v4.y, Tj = q4.y;
              if (KINDS & K_HEATPC) {                                      // (:124-129), in half-list orientation
                const int ff = P.fixflag[ij]; const double tc = P.tc[ij];
                double Ta = row_owns ? Ti : Tj, Tb = row_owns ? Tj : Ti;
                const int ta = row_owns ? ti : tj, tb = row_owns ? tj : ti;
                Ta = (ff == ta && Ta < Tb) ? tc : Ta;
                Tb = (ff == tb && Tb < Ta) ? tc : Tb;
                Ti = row_owns ? Ta : Tb; Tj = row_owns ? Tb : Ta;
              }
              const double term = fast_div15(2.0 * P.visc[ij] * (Ti - Tj) * wfd * mj, rhoi * rhoj);
              ade += hit ? term : 0.0;
            }
          }
        }
      }
      for (int o = lpw; o < 32; o <<= 1) {
        fx += __shfl_xor_sync(FULLMASK, fx, o); fy += __shfl_xor_sync(FULLMASK, fy, o); fz += __shfl_xor_sync(FULLMASK, fz, o);
        if (WRITES_DE) ade += __shfl_xor_sync(FULLMASK, ade, o);
      }
      if (valid && sub == 0) {
        double4 f = make_double4(0, 0, 0, 0); double de0 = 0.0;
        if (A.accum) { f = A.fd[dev]; if (WRITES_DE) de0 = A.de[dev]; }      // first force pass since force_clear: both are zero, no read-modify-write
        f.x += fx; f.y += fy; f.z += fz;
        A.fd[dev] = f;
        if (WRITES_DE) A.de[dev] = de0 + ade;
      }
    }
    L.release();
  }
}
