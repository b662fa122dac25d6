// b200_sph.cu -- host side of libb200sph.so: device state, the Verlet loop and the C-ABI
// of include/b200_sph.h.  sm_100a only; there is no CPU path in this library.
#include "b200_common.cuh"
#include "b200_neigh.cuh"
#include "b200_pair.cuh"
#include "b200_tile.cuh"
#include "b200_lj.cuh"
#include "b200_fix.cuh"
#include "b200_phase.cuh"
#include "b200_comm.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::string(#call) + ": " + cudaGetErrorString(e_); } while (0)
#define API_BEGIN try {
#define API_END } catch (const std::string &m) { return fail(m); } catch (const std::exception &e) { return fail(e.what()); } return 0;

enum { T_INTEGRATE = 0, T_COMM, T_NEIGH_BIN, T_NEIGH_BUILD, T_DENSITY, T_COLORGRAD, T_DERIVE, T_FORCE, T_FINAL, T_PHASE, T_NTIMERS };
static const char *timer_names[T_NTIMERS] = {"initial_integrate", "forward_comm", "neigh_bin_sort_ghost", "neigh_build", "density",
                                             "colorgradient", "records", "force", "reverse_post_final", "phase_change"};

template <class T> struct DevBuf {
  T *p = nullptr; size_t cap = 0;
  void ensure(size_t n, bool keep = false, cudaStream_t st = 0)
  {
    if (n <= cap) return;
    size_t nc = n + n / 8 + 64;
    T *q; CK(cudaMalloc(&q, nc * sizeof(T)));
    if (keep && p && cap) CK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (p) { CK(cudaStreamSynchronize(st)); CK(cudaFree(p)); }
    p = q; cap = nc;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct OwnedSet {
  DevBuf<double4> xt, vr, vm, fd, cgm;
  DevBuf<double> e, de, cv;
  DevBuf<int> tag, mask, orig;
  void ensure(size_t n, bool keep, cudaStream_t st)
  {
    xt.ensure(n, keep, st); vr.ensure(n, keep, st); vm.ensure(n, keep, st); fd.ensure(n, keep, st); cgm.ensure(n, keep, st);
    e.ensure(n, keep, st); de.ensure(n, keep, st); cv.ensure(n, keep, st);
    tag.ensure(n, keep, st); mask.ensure(n, keep, st); orig.ensure(n, keep, st);
  }
  void release() { xt.release(); vr.release(); vm.release(); fd.release(); cgm.release(); e.release(); de.release(); cv.release(); tag.release(); mask.release(); orig.release(); }
  OwnedArrays view() { return OwnedArrays{xt.p, vr.p, vm.p, fd.p, cgm.p, e.p, de.p, cv.p, tag.p, mask.p, orig.p}; }
};

// NCCL is bound at run time (dlopen) so that the library neither needs it on one GPU nor clashes with
// the copy a host framework (torch) already loaded; only ncclSend/ncclRecv/ncclAllReduce/ncclAllGather are used.
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  void load()
  {
    if (lib) return;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int k = 0; names[k] && !lib; k++) lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::string("b200: cannot dlopen libnccl.so.2 (needed for multi-GPU runs)");
#define BIND(f) *(void **)(&f) = dlsym(lib, "nccl" #f); if (!f) throw std::string("b200: libnccl lacks nccl" #f);
    BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(Send) BIND(Recv) BIND(AllReduce) BIND(AllGather) BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
  }
};
static NcclApi g_nccl;
#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) throw std::string(#call) + ": " + g_nccl.GetErrorString(r_); } while (0)

// one direction of the 6-way staged ghost exchange (comm_brick.cpp:330-386): who I send to / receive from,
// the slab of atoms I send, the periodic shift applied on sending
struct Swap {
  int dim, dir;                 // dir 0: send to the left neighbour (atoms near my lo face), 1: to the right
  int sendproc, recvproc;
  bool do_send, do_recv;
  double shift, slablo, slabhi;
  int imgstep;
  DevBuf<int> sendlist;
  int nsend = 0, nrecv = 0, firstrecv = 0;
  int last_nsend = -1, last_nrecv = -1;   // counts of the previous borders() (the caps of the next one), -1 = none
};

struct Pass { int type; int kinds; int nslots; int slots[4]; };
struct PcFix { b200_phase_change_desc d; long long next; int *d_state; };   // type: 0 rhosum 1 rhosum/mp 2 colorgradient 3 force

struct b200_sph {
  int device = 0;
  cudaStream_t st = 0;
  // problem
  Geom g{};
  bool have_domain = false, have_neigh = false;
  int multiphase = 0, ntypes = 0, ghost_velocity = 0;
  double mass[MAXT1] = {0};
  double skin = 0.3, cutneighmax = 0, triggersq = 0;
  int every = 1, delay = 10, check = 1, ago = 0;
  double h_cutneighsq[MAXTT] = {0}, h_farsq[MAXTT] = {0}, h_midsq[MAXTT] = {0};
  DevBuf<double> d_cutneighsq;
  double dt = 0, ftm2v = 1; long long ntimestep = 0;
  int npair = 0; PairTab h_tab[MAXPAIR]; PairTab *d_tab[MAXPAIR] = {nullptr};
  std::vector<Pass> plan;
  FixList fl{};
  std::vector<PcFix> pcs, pcs_old;             // pcs_old: the fixes of the previous registration (b200_fix_clear) -- an unchanged fix phase_change keeps its state
  DevBuf<unsigned char> pc_flag; DevBuf<double> pc_thr, pc_dmass; DevBuf<int> pc_dev; DevBuf<PcNew> pc_new;
  int maxtag = 0;
  // particles
  int nlocal = 0, nghost = 0;
  OwnedSet S[2]; int cur = 0;
  DevBuf<double4> rec;
  DevBuf<int> gimage;
  DevBuf<int> cellid, perm, perm2, gcell, gperm, gorder, flag, pos, alive;
  DevBuf<double> sendbuf, recvbuf, xs[2], xr[2];      // xs / xr: exact-size border messages of a swap without (or beyond) its cap
  // domain decomposition (one engine instance per rank / GPU)
  int world = 1, rank = 0, procgrid[3] = {1, 1, 1}, myloc[3] = {0, 0, 0}, procneigh[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  ncclComm_t nccl = nullptr;
  Swap swaps[6]; int nswap = 0;
  int next_orig = 0;
  // atom_modify sort (Atom::sortfreq / userbinsize / nextsort, atom.cpp:63-65): when the reference re-numbers its local indices
  int sortfreq = 0; double sort_binsize = 0.0; long long nextsort = 0; SortGeom sortgeom{}; bool sortgeom_ok = false;
  DevBuf<int> sort_cnt, sort_fill; bool sort_pending = false;
  double *d_red = nullptr, *h_red = nullptr;   // [0..7] scalars, [8..13] atom extent (-min, max per dimension), d_red[16..21] its ordered keys
  // boundary s / m (Domain::boundary, small, minxlo..): the box is re-fitted to the owned atoms on every rebuild
  int boundary[3][2] = {{0, 0}, {0, 0}, {0, 0}}; bool shrink = false; double small[3] = {0, 0, 0}, minbox[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  DevBuf<unsigned long long> key, gkey;
  DevBuf<int> cso, csg, cellfill, scan_tmp;
  DevBuf<double> xhold, stage_d, d_mass;
  DevBuf<int> stage_i;
  DevBuf<unsigned> nbr, far; DevBuf<int> numneigh, numfar; int stride = 32;
  DevBuf<double> d_prunesq, d_farsq, d_midsq; double far_margin = 0.0, mid_margin = 0.0;
  unsigned long long *d_dmaxsq = nullptr; int *d_scan_far = nullptr;
  // per-cell displacement bound -> per-tile zone flags (single-phase tile path; B200_ZONE_GLOBAL=1 keeps the one global flag)
  DevBuf<unsigned> celld; DevBuf<int> rowcell; DevBuf<unsigned char> tzone; bool zone_local = !getenv("B200_ZONE_GLOBAL");
  int *d_flags = nullptr, *h_flags = nullptr;   // [0] maxcount [1] moved flag [2] scratch
  // pair virial (Pair::virial_fdotr_compute) on request: the next force evaluation of b200_setup / the last step of b200_run
  bool vir_request = false, vir_now = false; DevBuf<double> virow, virpart; double *h_vir = nullptr;
  // fix dt/reset: device-resident timestep  d_dt[0] = dt, ((unsigned long long *)d_dt)[1] = running minimum (bits)
  bool dtreset = false; int dtr_bit = 0, dtr_every = 1, dtr_minbound = 0, dtr_maxbound = 0; double dtr_tmin = 0, dtr_tmax = 0, dtr_xmax = 0;
  double *d_dt = nullptr, *h_dtv = nullptr;      // h_dtv: pinned mirror of d_dt[0..4]
  double atime = 0.0; long long atimestep = 0, laststep = 0;      // Update::atime / atimestep, FixDtReset::laststep (b200_set_time / b200_get_time)
  // tile path (b200_tile.cuh): single-phase decks
  bool half_bin = false;      // no full-list sub-style in the deck: the reference's half list is half_bin_newton's (k_build, BuildArgs::hbn)
  bool tile_on = false, tile_ok = true, rows_tiled = false, tile_nouni = getenv("B200_TILE_NOUNI") != nullptr;
  int tile_nparts = 2, tile_nk = 1, tile_slotcap = 0, tile_cap = 0, ntiles = 0, nsm = 0, tile_split = 2;
  DevBuf<int> x_mark, x_inv, x_lflag, x_lpos, x_list, x_work;      // comm_exchange: replay of the reference's hole-filling order
  DevBuf<TileDesc> tiles, gtiles; DevBuf<double2> trec; DevBuf<int> rowtile;
  int ngtiles = 0;                              // tiles of ghost rows (multiphase styles)
  // halo overlap (single-phase tile path): tiles [0, nint) neither read ghosts nor feed a send list and run while the halo flies
  int nint = 0; bool comm_pending = false, no_overlap = getenv("B200_OVERLAP") == nullptr;     // measured slower than the plain order at C2 scale: opt-in
  cudaStream_t st2 = 0; cudaEvent_t ev_main = 0, ev_comm = 0;
  int *d_tflags = nullptr;                      // [0] ntiles [1] max slots [2] overflow [3] max rows [4] work counter; [8..11] the same for the ghost-row tiles
  bool setup_done = false, geom_ready = false;
  bool f_clean = false;                         // f, drho, de are all zero (force_clear just ran): the first tile force pass stores instead of adding
  // instrumentation
  long long launches = 0, nbuilds = 0, nsteps = 0, maxneigh = 0, ndanger = 0, ninserted = 0;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool; std::vector<int> ev_timer; size_t ev_used = 0;
  double t_ms[T_NTIMERS] = {0}; long long t_calls[T_NTIMERS] = {0};

  OwnedSet &C() { return S[cur]; }
  int nall() const { return nlocal + nghost; }
  void ensure_cap(size_t n, bool keep)
  {
    S[0].ensure(n, keep && cur == 0, st); S[1].ensure(n, keep && cur == 1, st);
    rec.ensure(n * 4, false, st); gimage.ensure(n, true, st);
    numneigh.ensure(n, false, st); numfar.ensure(n, false, st);
  }
  StepArrays step_arrays() { OwnedSet &c = C(); return StepArrays{c.xt.p, c.vr.p, c.vm.p, c.fd.p, c.e.p, c.de.p, c.mask.p, c.cv.p, c.tag.p}; }
  std::vector<ExprProg> progs; DevBuf<ExprProg> d_progs;       // compiled variable formulas of the registered fixes
  CommArrays comm_arrays() { OwnedSet &c = C(); return CommArrays{c.xt.p, c.vr.p, c.vm.p, c.fd.p, c.cgm.p, c.e.p, c.de.p, c.cv.p, c.tag.p, c.mask.p, c.orig.p, gimage.p}; }

  // ---- timing helpers ----
  void tbegin(int which)
  {
    if (!timing) return;
    if (ev_used == ev_pool.size()) { cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); ev_pool.push_back({a, b}); ev_timer.push_back(0); }
    ev_timer[ev_used] = which;
    CK(cudaEventRecord(ev_pool[ev_used].first, st));
  }
  void tend()
  {
    if (!timing) return;
    CK(cudaEventRecord(ev_pool[ev_used].second, st));
    ev_used++;
    if (ev_used >= 4096) tflush();
  }
  void tflush()
  {
    if (!ev_used) return;
    CK(cudaStreamSynchronize(st));
    for (size_t k = 0; k < ev_used; k++) {
      float ms = 0; CK(cudaEventElapsedTime(&ms, ev_pool[k].first, ev_pool[k].second));
      t_ms[ev_timer[k]] += ms; t_calls[ev_timer[k]]++;
    }
    ev_used = 0;
  }
};

static const bool g_sync_debug = getenv("B200_SYNC_DEBUG") != nullptr;   // synchronise + check after every launch
static void post_launch(b200_sph *h, const char *name)
{
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && g_sync_debug) e = cudaStreamSynchronize(h->st);
  if (e != cudaSuccess) throw std::string(name) + ": " + cudaGetErrorString(e);
}
#define LAUNCH(h, kern, grid, block, ...) do { kern<<<(grid), (block), 0, (h)->st>>>(__VA_ARGS__); post_launch(h, #kern); } while (0)
static inline int nblk(long long n, int b) { return (int)std::max<long long>(1, (n + b - 1) / b); }

// ------------------------------------------------------------------ scan ----
static void scan_exclusive(b200_sph *h, int *data, int n, int *scratch)
{
  // data[0..n) counts -> exclusive offsets, data[n] = total; recursive on the per-block totals
  if (n <= 0) { CK(cudaMemsetAsync(data, 0, sizeof(int), h->st)); return; }      // an empty rank (no owned atoms, no slots): total 0
  int per = SCAN_T * SCAN_E, nb = (n + per - 1) / per;
  LAUNCH(h, k_scan_block, nb, SCAN_T, data, n, scratch);
  LAUNCH(h, k_scan_sums, 1, 1024, scratch, nb, data + n);
  if (nb > 1) LAUNCH(h, k_scan_add, nb, SCAN_T, data, n, scratch);
}

static bool zones_on(const b200_sph *h) { return h->zone_local && h->far_margin > 0.0 && h->rows_tiled && !h->multiphase && h->ntiles > 0; }

// -------------------------------------------------------------- geometry ----
static void setup_geometry(b200_sph *h, bool keep_comm_history = false)
{
  Geom &g = h->g;
  if (!h->have_domain || !h->have_neigh) throw std::string("b200_domain / b200_neighbor must be called before setup");
  double cut = h->cutneighmax;
  for (int d = 0; d < 3; d++) {
    g.prd[d] = g.boxhi[d] - g.boxlo[d];
    g.slab_lo_hi[d] = g.sublo[d] + g.cutghost;
    g.slab_hi_lo[d] = g.subhi[d] - g.cutghost;
    bool swaps = (g.periodic[d] || h->procgrid[d] > 1) && !(g.dim == 2 && d == 2);
    if (swaps) {
      // one ghost layer of neighbour ranks: maxneed = 1 (comm_brick.cpp:225-231 for uniform bricks, the updown() walk :260-300 for
      // non-uniform ones, i.e. after `balance ... shift`): every brick must be at least one ghost cutoff long
      int maxneed = (int)(g.cutghost * h->procgrid[d] / g.prd[d]) + 1;
      if (h->procgrid[d] > 1 && g.subhi[d] - g.sublo[d] < g.cutghost) maxneed = 2;
      if (maxneed > 1) throw std::string("b200: ghost cutoff >= sub-domain length is not supported (one ghost layer of neighbour ranks)");
    }
    double lo = swaps ? g.sublo[d] - g.cutghost : g.sublo[d];
    double hi = swaps ? g.subhi[d] + g.cutghost : g.subhi[d];
    double len = hi - lo;
    int nc = (cut > 0.0) ? (int)(len / cut) : 1;
    if (nc < 1) nc = 1;
    if (g.dim == 2 && d == 2) nc = 1;
    if (nc > 4000) nc = 4000;
    g.clo[d] = lo; g.nc[d] = nc; g.cinv[d] = nc / len;
  }
  g.ncells = g.nc[0] * g.nc[1] * g.nc[2];
  // the reference's bins: Neighbor::setup_bins, neighbor.cpp:1618-1735
  double binsize_optimal = 0.5 * cut;
  if (binsize_optimal == 0.0) binsize_optimal = g.prd[0];
  double binsizeinv = 1.0 / binsize_optimal;
  for (int d = 0; d < 3; d++) {
    int nb = (int)(g.prd[d] * binsizeinv);
    if (g.dim == 2 && d == 2) nb = 1;
    if (nb == 0) nb = 1;
    g.nbin[d] = nb; g.binsize[d] = g.prd[d] / nb; g.bininv[d] = 1.0 / g.binsize[d];
    if (nb + 64 > RB_BIAS * 32) throw std::string("b200: too many reference bins per dimension");
  }
  int s[3];
  for (int d = 0; d < 3; d++) { s[d] = (int)(cut * g.bininv[d]); if (s[d] * g.binsize[d] < cut) s[d]++; }
  if (g.dim == 2) s[2] = 0;
  g.sx = s[0]; g.sy = s[1]; g.sz = s[2];
  g.cutneighmaxsq = cut * cut;
  h->cso.ensure(g.ncells + 2); h->csg.ensure(g.ncells + 2); h->cellfill.ensure(g.ncells + 2);
  CK(cudaMemsetAsync(h->csg.p, 0, (g.ncells + 2) * sizeof(int), h->st));
  // swaps: CommBrick::setup, comm_brick.cpp:330-386 (maxneed = 1)
  h->nswap = 0;
  for (int d = 0; d < 3; d++) {
    bool any = (g.periodic[d] || h->procgrid[d] > 1) && !(g.dim == 2 && d == 2);
    if (!any) continue;
    for (int dir = 0; dir < 2; dir++) {
      Swap &s = h->swaps[h->nswap++];
      s.dim = d; s.dir = dir;
      s.sendproc = h->procneigh[d][dir]; s.recvproc = h->procneigh[d][dir ^ 1];
      int loc = h->myloc[d], pg = h->procgrid[d];
      if (dir == 0) {
        s.slablo = -1.0e300; s.slabhi = g.sublo[d] + g.cutghost;
        s.do_send = g.periodic[d] || loc > 0; s.do_recv = g.periodic[d] || loc < pg - 1;
        s.shift = (loc == 0) ? g.prd[d] : 0.0; s.imgstep = (loc == 0) ? 1 : 0;
      } else {
        s.slablo = g.subhi[d] - g.cutghost; s.slabhi = 1.0e300;
        s.do_send = g.periodic[d] || loc < pg - 1; s.do_recv = g.periodic[d] || loc > 0;
        s.shift = (loc == pg - 1) ? -g.prd[d] : 0.0; s.imgstep = (loc == pg - 1) ? -1 : 0;
      }
      s.imgstep *= (d == 0 ? 1 : (d == 1 ? 3 : 9));
      s.nsend = s.nrecv = 0;
      if (!keep_comm_history) s.last_nsend = s.last_nrecv = -1;      // a new domain / decomposition: the swaps' message caps start over (on every rank alike)
    }
  }
  h->geom_ready = true;
}

// Domain::reset_box (domain.cpp:338-406): shrink-wrapped faces follow the extent of the owned atoms of ALL ranks; then set_global_box /
// set_local_box (:268-330, uniform xsplit) and everything derived from the box: comm->setup, neighbor->setup_bins (verlet.cpp:102, 244-248)
static void refit_box(b200_sph *h)
{
  Geom &g = h->g;
  unsigned long long *keys = (unsigned long long *)(h->d_red + 16);
  LAUNCH(h, k_extent_init, 1, 32, keys);
  if (h->nlocal) LAUNCH(h, k_extent, nblk(h->nlocal, 256), 256, h->nlocal, h->C().xt.p, keys);
  LAUNCH(h, k_extent_decode, 1, 32, keys, h->d_red + 8);
  if (h->world > 1) NCK(g_nccl.AllReduce(h->d_red + 8, h->d_red + 8, 6, ncclDouble, ncclMax, h->nccl, h->st));
  CK(cudaMemcpyAsync(h->h_red + 8, h->d_red + 8, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  const double *all = h->h_red + 8;
  for (int d = 0; d < 3; d++) {
    if (g.periodic[d]) continue;
    if (h->boundary[d][0] == 2) g.boxlo[d] = -all[2 * d] - h->small[d];
    else if (h->boundary[d][0] == 3) g.boxlo[d] = std::min(-all[2 * d] - h->small[d], h->minbox[d][0]);
    if (h->boundary[d][1] == 2) g.boxhi[d] = all[2 * d + 1] + h->small[d];
    else if (h->boundary[d][1] == 3) g.boxhi[d] = std::max(all[2 * d + 1] + h->small[d], h->minbox[d][1]);
    if (g.boxlo[d] > g.boxhi[d]) throw std::string("Illegal simulation box");
  }
  for (int d = 0; d < 3; d++) {
    double prd = g.boxhi[d] - g.boxlo[d];
    int loc = h->myloc[d], pg = h->procgrid[d];
    g.sublo[d] = g.boxlo[d] + prd * (loc * 1.0 / pg);
    g.subhi[d] = loc < pg - 1 ? g.boxlo[d] + prd * ((loc + 1) * 1.0 / pg) : g.boxhi[d];
  }
  setup_geometry(h, true);
}

// ------------------------------------------------------------------ comm ----
static void scan_exclusive(b200_sph *h, int *data, int n, int *scratch);
static void ensure_scan_tmp(b200_sph *h, size_t n) { h->scan_tmp.ensure(n / (SCAN_T * SCAN_E) * 2 + 4096); }
static bool swap_self(const b200_sph *h, const Swap &s) { return s.sendproc == h->rank && s.recvproc == h->rank; }   // "if swapping with self, simply copy" (comm_brick.cpp:800)

// The two swaps of one dimension are independent of each other (both scan the atoms present before the dimension, comm_brick.cpp:722-725),
// so they travel together: swaps [k0, k1) pack into consecutive regions of sendbuf, ONE grouped NCCL call moves both directions,
// then both are unpacked.  nsend_d / nrecv_d: doubles per swap.  Returns the buffers the receivers unpack from (rbuf[k - k0]).
static void dim_transfer(b200_sph *h, int k0, int k1, const size_t *nsend_d, const size_t *nrecv_d, const size_t *soff, const size_t *roff, double **rbuf, bool reverse)
{
  bool any = false;
  for (int k = k0; k < k1; k++) {
    Swap &s = h->swaps[k];
    if (swap_self(h, s)) { rbuf[k - k0] = h->sendbuf.p + soff[k - k0]; continue; }
    rbuf[k - k0] = h->recvbuf.p + roff[k - k0];
    if (nsend_d[k - k0] || nrecv_d[k - k0]) any = true;
  }
  if (!any) return;
  NCK(g_nccl.GroupStart());
  for (int k = k0; k < k1; k++) {
    Swap &s = h->swaps[k];
    if (swap_self(h, s)) continue;
    // forward: my slab goes to sendproc, ghosts come from recvproc; reverse: the ghosts' sums go back to recvproc
    if (nsend_d[k - k0]) NCK(g_nccl.Send(h->sendbuf.p + soff[k - k0], nsend_d[k - k0], ncclDouble, reverse ? s.recvproc : s.sendproc, h->nccl, h->st));
    if (nrecv_d[k - k0]) NCK(g_nccl.Recv(h->recvbuf.p + roff[k - k0], nrecv_d[k - k0], ncclDouble, reverse ? s.sendproc : s.recvproc, h->nccl, h->st));
  }
  NCK(g_nccl.GroupEnd());
}
static int dim_end(const b200_sph *h, int k0) { int k = k0; while (k < h->nswap && h->swaps[k].dim == h->swaps[k0].dim) k++; return k; }

// CommBrick::borders, comm_brick.cpp:696-864.  One host synchronisation per DIMENSION: the counts travel in the message headers
// (k_pack_border_compact) and the messages are sized by a cap both sides derive from the previous build's count of the same swap;
// a swap without history (first build) or whose count outgrew the cap takes the exact two-step exchange (count, then data).
static void comm_borders(b200_sph *h)
{
  const int B = 256;
  Geom &g = h->g;
  h->nghost = 0;
  for (int k0 = 0; k0 < h->nswap; k0 = dim_end(h, k0)) {
    const int k1 = dim_end(h, k0), dim = h->swaps[k0].dim, ns = k1 - k0;
    const int nlast = h->nlocal + h->nghost;                 // both swaps of a dimension scan the atoms present before it (:722-725)
    int *flag[2], *pos[2]; int capS[2], capR[2]; size_t soff[2], roff[2], nsd[2], nrd[2]; double *rbuf[2];
    if (ns != 2) throw std::string("b200: a dimension with other than two swaps (more than one ghost layer)");
    h->flag.ensure((size_t)2 * nlast + 4); h->pos.ensure((size_t)2 * nlast + 4); ensure_scan_tmp(h, 2 * (size_t)nlast + 4);
    size_t so = 0, ro = 0;
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      flag[q] = h->flag.p + (size_t)q * nlast; pos[q] = h->pos.p + (size_t)q * nlast;      // one scan serves both swaps (k_slab_flag2)
      // cap from the last count of this swap: the sender's last nsend IS the receiver's last nrecv, so both sides size alike
      capS[q] = (s.do_send && s.last_nsend >= 0) ? s.last_nsend + s.last_nsend / 4 + 64 : -1;
      capR[q] = (s.do_recv && s.last_nrecv >= 0) ? s.last_nrecv + s.last_nrecv / 4 + 64 : -1;
      if (swap_self(h, s)) capR[q] = capS[q];
      soff[q] = so; roff[q] = ro;
      so += (size_t)std::max(capS[q], 0) * NB_BORDER + 1; ro += (size_t)std::max(capR[q], 0) * NB_BORDER + 1;
      s.nsend = s.nrecv = 0;
    }
    h->sendbuf.ensure(so + 1); h->recvbuf.ensure(ro + 1);
    // senders: flag both swaps in one pass, one scan (+ compacting pack where a cap exists)
    const int *bias[2] = {nullptr, h->pos.p + nlast};             // swap 1's offsets start at count0 = P[nlast]
    if (nlast) {
      Swap &s0 = h->swaps[k0], &s1 = h->swaps[k0 + 1];
      LAUNCH(h, k_slab_flag2, nblk(nlast, B), B, nlast, h->C().xt.p, dim, s0.slablo, s0.slabhi, s0.do_send ? 1 : 0, s1.slablo, s1.slabhi, s1.do_send ? 1 : 0,
             h->flag.p, h->pos.p);
      scan_exclusive(h, h->pos.p, 2 * nlast, h->scan_tmp.p);
    } else CK(cudaMemsetAsync(h->pos.p, 0, 2 * sizeof(int), h->st));
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      if (!s.do_send) continue;
      if (capS[q] >= 0) {
        s.sendlist.ensure(capS[q] + 1);
        LAUNCH(h, k_pack_border_compact, nblk(std::max(nlast, 1), B), B, nlast, flag[q], pos[q], bias[q], capS[q], s.sendlist.p, h->comm_arrays(), dim, s.shift, s.imgstep,
               k0 + q + 1, h->sendbuf.p + soff[q]);
      }
    }
    // P[nlast] = count0, P[2 nlast] = count0 + count1 (a swap that does not send flagged nothing)
    CK(cudaMemcpyAsync(h->h_flags + 28, h->pos.p + nlast, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->h_flags + 29, h->pos.p + 2 * (size_t)nlast, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    // fast path: header + cap records in one grouped call for the swaps that have history on both ends
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      nsd[q] = (s.do_send && capS[q] >= 0) ? (size_t)capS[q] * NB_BORDER + 1 : 0;
      nrd[q] = (s.do_recv && capR[q] >= 0 && !swap_self(h, s)) ? (size_t)capR[q] * NB_BORDER + 1 : 0;
    }
    dim_transfer(h, k0, k1, nsd, nrd, soff, roff, rbuf, false);
    for (int q = 0; q < ns; q++)
      if (h->swaps[k0 + q].do_recv && capR[q] >= 0) CK(cudaMemcpyAsync(h->h_red + 14 + q, rbuf[q], sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));                       // the one synchronisation of this dimension
    bool redo[2] = {false, false};
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      s.nsend = s.do_send ? (q == 0 ? h->h_flags[28] : h->h_flags[29] - h->h_flags[28]) : 0;
      if (s.do_send && (capS[q] < 0 || s.nsend > capS[q])) redo[q] = true;
      if (s.do_recv && capR[q] >= 0) { s.nrecv = (int)h->h_red[14 + q]; if (s.nrecv > capR[q]) redo[q] = true; }
      else if (s.do_recv) redo[q] = true;
      if (swap_self(h, s)) s.nrecv = s.nsend;
    }
    // exact path for the swaps that need it.  Pairwise consistent: the sender of a swap and its receiver hold the same cap (same
    // history) and see the same count (header), so "send exact" on one end <=> "receive exact" on the other; every rank walks
    // q = 0, 1 in the same order and posts a swap's operations in one group.
    for (int q = 0; q < ns; q++) {
      if (!redo[q]) continue;
      Swap &s = h->swaps[k0 + q];
      const bool self = swap_self(h, s);
      const bool send_exact = s.do_send && (capS[q] < 0 || s.nsend > capS[q]);
      if (send_exact) {
        s.sendlist.ensure(s.nsend + 1);
        h->xs[q].ensure((size_t)s.nsend * NB_BORDER + 2);
        if (s.nsend) {
          LAUNCH(h, k_compact, nblk(nlast, B), B, nlast, flag[q], pos[q], s.sendlist.p, bias[q]);
          LAUNCH(h, k_pack_border, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, h->comm_arrays(), dim, s.shift, s.imgstep, k0 + q + 1, h->xs[q].p + 1);
        }
      }
      if (self) { s.nrecv = s.nsend; if (send_exact) rbuf[q] = h->xs[q].p; continue; }
      const bool count_send = s.do_send && capS[q] < 0, count_recv = s.do_recv && capR[q] < 0;
      if (count_send || count_recv) {      // no history: the 1-int MPI_Sendrecv of borders() (:819)
        h->h_red[q] = (double)s.nsend;                        // staging slots per swap: the async copy of q = 0 may still be pending at q = 1
        CK(cudaMemcpyAsync(h->d_red + q, h->h_red + q, sizeof(double), cudaMemcpyHostToDevice, h->st));
        NCK(g_nccl.GroupStart());
        if (count_send) NCK(g_nccl.Send(h->d_red + q, 1, ncclDouble, s.sendproc, h->nccl, h->st));
        if (count_recv) NCK(g_nccl.Recv(h->d_red + 2 + q, 1, ncclDouble, s.recvproc, h->nccl, h->st));
        NCK(g_nccl.GroupEnd());
        if (count_recv) {
          CK(cudaMemcpyAsync(h->h_red + 2 + q, h->d_red + 2 + q, sizeof(double), cudaMemcpyDeviceToHost, h->st));
          CK(cudaStreamSynchronize(h->st));
          s.nrecv = (int)h->h_red[2 + q];
        }
      }
      const bool recv_exact = s.do_recv && (capR[q] < 0 || s.nrecv > capR[q]);
      if (recv_exact) { h->xr[q].ensure((size_t)s.nrecv * NB_BORDER + 2); rbuf[q] = h->xr[q].p; }
      NCK(g_nccl.GroupStart());
      if (send_exact && s.nsend) NCK(g_nccl.Send(h->xs[q].p + 1, (size_t)s.nsend * NB_BORDER, ncclDouble, s.sendproc, h->nccl, h->st));
      if (recv_exact && s.nrecv) NCK(g_nccl.Recv(h->xr[q].p + 1, (size_t)s.nrecv * NB_BORDER, ncclDouble, s.recvproc, h->nccl, h->st));
      NCK(g_nccl.GroupEnd());
    }
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      if (!s.do_recv) s.nrecv = 0;
      s.firstrecv = h->nlocal + h->nghost;
      h->ensure_cap((size_t)s.firstrecv + s.nrecv, true);
      if (s.nrecv) LAUNCH(h, k_unpack_border, nblk(s.nrecv, B), B, g, s.nrecv, s.firstrecv, h->comm_arrays(), rbuf[q] + 1);
      h->nghost += s.nrecv;
      s.last_nsend = s.do_send ? s.nsend : -1; s.last_nrecv = s.do_recv ? s.nrecv : -1;
    }
  }
}
// generic staged swap loop, one grouped transfer per dimension: forward direction
template <class Pack, class Unpack> static void comm_forward_generic(b200_sph *h, int width, Pack pack, Unpack unpack)
{
  for (int k0 = 0; k0 < h->nswap; k0 = dim_end(h, k0)) {
    const int k1 = dim_end(h, k0), ns = k1 - k0;
    size_t soff[2], roff[2], nsd[2], nrd[2]; double *rbuf[2];
    size_t so = 0, ro = 0;
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      nsd[q] = (size_t)s.nsend * width; nrd[q] = swap_self(h, s) ? 0 : (size_t)s.nrecv * width;
      soff[q] = so; roff[q] = ro; so += nsd[q]; ro += nrd[q];
    }
    h->sendbuf.ensure(so + 1); h->recvbuf.ensure(ro + 1);
    for (int q = 0; q < ns; q++) if (h->swaps[k0 + q].nsend) pack(h->swaps[k0 + q], h->sendbuf.p + soff[q]);
    dim_transfer(h, k0, k1, nsd, nrd, soff, roff, rbuf, false);
    for (int q = 0; q < ns; q++) if (h->swaps[k0 + q].nrecv) unpack(h->swaps[k0 + q], rbuf[q]);
  }
}
// reverse direction: ghosts' values go back to the atoms they were copied from (comm_brick.cpp:513-560); dimensions and, inside a
// dimension, the unpacks run in the reference's reverse swap order (an atom in both send lists receives its two additions in that order)
template <class Pack, class Unpack> static void comm_reverse_generic(b200_sph *h, int width, Pack pack, Unpack unpack)
{
  std::vector<int> starts;
  for (int k0 = 0; k0 < h->nswap; k0 = dim_end(h, k0)) starts.push_back(k0);
  for (int d = (int)starts.size() - 1; d >= 0; d--) {
    const int k0 = starts[d], k1 = dim_end(h, k0), ns = k1 - k0;
    size_t soff[2], roff[2], nsd[2], nrd[2]; double *rbuf[2];
    size_t so = 0, ro = 0;
    for (int q = 0; q < ns; q++) {
      Swap &s = h->swaps[k0 + q];
      nsd[q] = (size_t)s.nrecv * width; nrd[q] = swap_self(h, s) ? 0 : (size_t)s.nsend * width;
      soff[q] = so; roff[q] = ro; so += nsd[q]; ro += nrd[q];
    }
    h->sendbuf.ensure(so + 1); h->recvbuf.ensure(ro + 1);
    for (int q = ns - 1; q >= 0; q--) if (h->swaps[k0 + q].nrecv) pack(h->swaps[k0 + q], h->sendbuf.p + soff[q]);
    dim_transfer(h, k0, k1, nsd, nrd, soff, roff, rbuf, true);
    for (int q = ns - 1; q >= 0; q--) if (h->swaps[k0 + q].nsend) unpack(h->swaps[k0 + q], rbuf[q]);
  }
}
static void comm_reverse_scalar_add(b200_sph *h, double *arr)      // comm->reverse_comm_fix (dmass)
{
  const int B = 256;
  comm_reverse_generic(h, 1,
    [&](Swap &s, double *dst) { LAUNCH(h, k_pack_scalar, nblk(s.nrecv, B), B, s.nrecv, s.firstrecv, arr, dst); },
    [&](Swap &s, double *buf) { LAUNCH(h, k_unpack_scalar_add, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, arr, buf); });
}
// CommBrick::exchange, comm_brick.cpp:573-684: returns the number of slots now in use (dead ones included)
static int comm_exchange(b200_sph *h, int nslots)
{
  const int B = 256;
  Geom &g = h->g;
  for (int d = 0; d < 3; d++) {
    if (h->procgrid[d] == 1 || (g.dim == 2 && d == 2)) continue;
    h->flag.ensure(nslots + 2); h->pos.ensure(nslots + 2); ensure_scan_tmp(h, nslots + 2);
    LAUNCH(h, k_exchange_flag, nblk(nslots, B), B, nslots, h->C().xt.p, h->alive.p, d, g.sublo[d], g.subhi[d], h->flag.p, h->pos.p);
    scan_exclusive(h, h->pos.p, nslots, h->scan_tmp.p);
    CK(cudaMemcpyAsync(h->h_flags + 4, h->pos.p + nslots, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    int nsend = h->h_flags[4];
    h->perm.ensure(nsend + 1);
    h->sendbuf.ensure((size_t)nsend * NB_EXCHANGE + 1);
    if (nsend) {
      // the reference's hole-filling walk on the local-index keys (b200_comm.cuh): dense indices, leavers by index, the chain, the pack order
      const int M = std::max(h->next_orig, nslots) + 1;
      h->x_mark.ensure((size_t)M + 2); h->x_inv.ensure((size_t)nslots + 2); h->x_lflag.ensure((size_t)nslots + 2); h->x_lpos.ensure((size_t)nslots + 2);
      h->x_list.ensure((size_t)nsend + 2); ensure_scan_tmp(h, (size_t)M + 2);
      CK(cudaMemsetAsync(h->x_mark.p, 0, (size_t)(M + 1) * sizeof(int), h->st));
      LAUNCH(h, k_orig_mark, nblk(nslots, B), B, nslots, h->alive.p, h->C().orig.p, h->x_mark.p);
      scan_exclusive(h, h->x_mark.p, M, h->scan_tmp.p);
      CK(cudaMemsetAsync(h->x_lflag.p, 0, (size_t)(nslots + 1) * sizeof(int), h->st));      // indices >= the number of live atoms hold no leaver
      LAUNCH(h, k_orig_rank, nblk(nslots, B), B, nslots, h->alive.p, h->C().orig.p, h->x_mark.p, h->x_inv.p, h->flag.p, h->x_lflag.p);
      CK(cudaMemcpyAsync(h->x_lpos.p, h->x_lflag.p, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToDevice, h->st));
      scan_exclusive(h, h->x_lpos.p, nslots, h->scan_tmp.p);
      LAUNCH(h, k_index_compact, nblk(nslots, B), B, nslots, h->x_lflag.p, h->x_lpos.p, h->x_list.p);
      h->x_work.ensure((size_t)3 * nsend + 4);
      LAUNCH(h, k_holefill, 1, 256, h->x_list.p, nsend, h->x_mark.p + M, h->x_inv.p, h->flag.p, h->C().orig.p, h->perm.p, h->x_work.p);      // x_mark[M] = live atoms (no host round trip)
      h->next_orig = nslots - nsend;              // above every staying atom's index (they hold 0 .. live - nsend - 1 again); arrivals continue from here
      LAUNCH(h, k_pack_exchange, nblk(nsend, B), B, nsend, h->perm.p, h->comm_arrays(), h->alive.p, h->sendbuf.p);
    }
    // the whole buffer goes to both neighbours; each keeps what falls inside its bounds (:640-664)
    int left = h->procneigh[d][0], right = h->procneigh[d][1];
    int nrecv[2] = {0, 0};
    for (int side = 0; side < (left == right ? 1 : 2); side++) {
      int to = side ? right : left, from = side ? left : right;
      h->h_red[0] = (double)nsend;
      CK(cudaMemcpyAsync(h->d_red, h->h_red, sizeof(double), cudaMemcpyHostToDevice, h->st));
      NCK(g_nccl.GroupStart());
      NCK(g_nccl.Send(h->d_red, 1, ncclDouble, to, h->nccl, h->st));
      NCK(g_nccl.Recv(h->d_red + 1, 1, ncclDouble, from, h->nccl, h->st));
      NCK(g_nccl.GroupEnd());
      CK(cudaMemcpyAsync(h->h_red + 1, h->d_red + 1, sizeof(double), cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      nrecv[side] = (int)h->h_red[1];
      h->recvbuf.ensure((size_t)nrecv[side] * NB_EXCHANGE + 1);
      NCK(g_nccl.GroupStart());
      if (nsend) NCK(g_nccl.Send(h->sendbuf.p, (size_t)nsend * NB_EXCHANGE, ncclDouble, to, h->nccl, h->st));
      if (nrecv[side]) NCK(g_nccl.Recv(h->recvbuf.p, (size_t)nrecv[side] * NB_EXCHANGE, ncclDouble, from, h->nccl, h->st));
      NCK(g_nccl.GroupEnd());
      if (nrecv[side]) {
        h->ensure_cap((size_t)nslots + nrecv[side], true); h->alive.ensure((size_t)nslots + nrecv[side] + 1, true, h->st);
        CK(cudaMemsetAsync(h->d_flags + 3, 0, sizeof(int), h->st));
        LAUNCH(h, k_unpack_exchange, nblk(nrecv[side], B), B, nrecv[side], h->recvbuf.p, d, g.sublo[d], g.subhi[d], nslots, h->comm_arrays(), h->alive.p,
               h->d_flags + 3, h->next_orig);
        nslots += nrecv[side]; h->next_orig += nrecv[side];
      }
    }
  }
  return nslots;
}


// ------------------------------------------------------------- tile path ----
// persistent launch of a tile kernel: as many CTAs as fit on the device (or tiles), dynamic shared memory opted in once per kernel
// opt a tile kernel into the 227 KB of dynamic shared memory (once per kernel); returns what it may use
static size_t tile_optin(const void *fp)
{
  static std::vector<std::pair<const void *, size_t>> opted;     // kernel -> dynamic shared memory it may use
  for (auto &o : opted) if (o.first == fp) return o.second;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, fp));
  size_t maxdyn = TILE_SMEM_MAX - fa.sharedSizeBytes;
  CK(cudaFuncSetAttribute(fp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxdyn));
  opted.push_back({fp, maxdyn});
  return maxdyn;
}
template <class K, class A> static void launch_tiles(b200_sph *h, K kern, const char *name, int nthreads, size_t smem, const A &args, int ntiles = -1, int reserve_sms = 0)
{
  if (ntiles < 0) ntiles = h->ntiles;
  if (!ntiles) return;
  const void *fp = (const void *)kern;
  const size_t maxdyn = tile_optin(fp);
  if (smem > maxdyn) throw std::string(name) + ": tile does not fit in shared memory";
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fp, nthreads, smem));
  if (occ < 1) throw std::string(name) + ": zero occupancy";
  int grid = std::max(1, std::min(ntiles, (h->nsm - reserve_sms) * occ));     // reserve_sms: leave room for the halo's NCCL / pack kernels
  CK(cudaMemsetAsync(h->d_tflags + 4, 0, sizeof(int), h->st));
  kern<<<grid, nthreads, smem, h->st>>>(args);
  post_launch(h, name);
}

// plan + rows of the tile path; false = a single cell's neighbourhood does not fit in shared memory (row path takes over)
static bool tile_rows(b200_sph *h)
{
  Geom &g = h->g;
  int nl = h->nlocal, na = h->nall();
  const bool mp = h->multiphase != 0;
  h->ntiles = h->ngtiles = 0;
  CK(cudaMemsetAsync(h->d_tflags, 0, 16 * sizeof(int), h->st));
  bool classes = false;
  if (nl) {
    h->tiles.ensure((mp ? 2 : 1) * ((size_t)g.ncells + 1));      // multiphase: the ghost-row tiles are appended (below)
    TilePlanArgs P{g, nl, TILE_ROWS, h->tile_slotcap, 0, mp ? 1 : 0, -1, {0, 0, 0}, h->cso.p, h->csg.p, h->tiles.p, h->d_tflags};
    classes = !mp && h->nswap > 0 && !h->no_overlap;
    for (int k = 0; k < h->nswap; k++)       // the interior criterion of k_tile_plan needs cells at least one ghost cutoff wide
      if (1.0 / g.cinv[h->swaps[k].dim] < g.cutghost) classes = false;
    if (classes) {             // interior tiles first, then the boundary tiles (k_tile_plan)
      for (int k = 0; k < h->nswap; k++) P.swapdim[h->swaps[k].dim] = 1;
      P.want = 0;
      LAUNCH(h, k_tile_plan, nblk((long long)g.nc[1] * g.nc[2] * 32, 128), 128, P);
      CK(cudaMemcpyAsync(h->d_tflags + 5, h->d_tflags, sizeof(int), cudaMemcpyDeviceToDevice, h->st));
      P.want = 1;
    }
    LAUNCH(h, k_tile_plan, nblk((long long)g.nc[1] * g.nc[2] * 32, 128), 128, P);
    if (mp && h->nghost) {       // ghost rows: what the reference accumulates on ghost atoms (then reverse-communicates)
      h->gtiles.ensure((size_t)g.ncells + 1);
      TilePlanArgs G{g, nl, TILE_ROWS, h->tile_slotcap, 1, 1, -1, {0, 0, 0}, h->cso.p, h->csg.p, h->gtiles.p, h->d_tflags + 8};
      LAUNCH(h, k_tile_plan, nblk((long long)g.nc[1] * g.nc[2] * 32, 128), 128, G);
    }
  }
  // tiles or rows is a collective decision: the single-phase tile path sends no reverse halo, the row path does, so a rank
  // falling back alone would leave its peers waiting in ncclRecv
  if (h->world > 1) {
    NCK(g_nccl.AllReduce(h->d_tflags + 2, h->d_tflags + 2, 1, ncclInt, ncclMax, h->nccl, h->st));
    NCK(g_nccl.AllReduce(h->d_tflags + 10, h->d_tflags + 10, 1, ncclInt, ncclMax, h->nccl, h->st));
  }
  int *hf = h->h_flags + 16;                       // pinned: [0..3] owned-row tiles, [8..11] ghost-row tiles
  CK(cudaMemcpyAsync(hf, h->d_tflags, 12 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (hf[2] || hf[10]) return false;
  if (!nl) return true;
  h->ntiles = hf[0]; h->ngtiles = hf[8]; h->nint = classes ? hf[5] : 0;
  h->tile_cap = std::max(2, (std::max(hf[1], hf[9]) + 1) & ~1);
  // one tile queue: owned-row tiles [0, ntiles), then the ghost-row tiles -- the build and the force pass walk both in ONE persistent
  // launch (the short ghost tiles fill the tail of the owned ones; profiles/r02_launches_c3: 0.26 + 0.13 ms as launches of their own)
  if (h->ngtiles) CK(cudaMemcpyAsync(h->tiles.p + h->ntiles, h->gtiles.p, (size_t)h->ngtiles * sizeof(TileDesc), cudaMemcpyDeviceToDevice, h->st));
  const int nrows = mp ? na : nl;
  if (mp) h->rowtile.ensure(nl + 1);
  for (int attempt = 0; attempt < 8; attempt++) {
    size_t rows32 = (size_t)(nrows + 31) / 32 * 32;
    int ngrp = h->stride / 8;
    h->nbr.ensure(rows32 * ngrp * 4); h->far.ensure(rows32 * ngrp * 4);
    CK(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(int), h->st));
    CK(cudaMemsetAsync(h->d_flags + 11, 0, sizeof(int), h->st));
    TileBuildArgs B{};
    B.maxn = h->d_flags + 11;
    B.g = g; B.nlocal = nl; B.ngrp = ngrp; B.cap = h->tile_cap;
    B.xt = h->C().xt.p; B.gorder = h->gorder.p; B.cso = h->cso.p; B.csg = h->csg.p;
    B.cutneighsq = h->d_cutneighsq.p; B.farsq = h->d_farsq.p; B.midsq = h->d_midsq.p;
    B.tiles = h->tiles.p; B.ntiles = h->ntiles; B.counter = h->d_tflags + 4;
    B.near = (uint4 *)h->nbr.p; B.far = (uint4 *)h->far.p; B.numneigh = h->numneigh.p; B.numfar = h->numfar.p; B.maxcount = h->d_flags;
    B.orig = h->C().orig.p; B.rowtile = mp ? h->rowtile.p : nullptr;
    // one cutoff for every type pair (the common deck) -> scalar thresholds in the fp32 phase
    B.uni = 1; B.cutsq_u = h->h_cutneighsq[1 * MAXT1 + 1]; B.farsq_u = h->h_farsq[1 * MAXT1 + 1]; B.midsq_u = h->h_midsq[1 * MAXT1 + 1];
    for (int i = 1; i <= h->ntypes; i++) for (int j = 1; j <= h->ntypes; j++)
      if (h->h_cutneighsq[i * MAXT1 + j] != B.cutsq_u || h->h_farsq[i * MAXT1 + j] != B.farsq_u || h->h_midsq[i * MAXT1 + j] != B.midsq_u) B.uni = 0;
    size_t bsm = (size_t)(mp ? 17 : 13) * (((h->tile_cap + 3) & ~3) + 4);
    {
      const int nt = h->ntiles + h->ngtiles;
      B.ntiles = nt;
      bool small = bsm <= 40 * 1024;      // small tiles (the 27-row cells of the multiphase decks, 35 KB: 1.94 ms with 128 threads, 2.92 with 256): 128-thread CTAs, more of them
                                          // per SM; the C2 tiles (46 KB) build 5 % faster with 256 threads (2.11 vs 2.23 ms, gpurun r02 last A/B)
      if (const char *e = getenv("B200_BUILD_NT")) small = atoi(e) == 128;
      const bool zones = h->far_margin > 0.0 || !B.uni;      // skin 0: no far / mid rows to split off
#define BUILD_LAUNCH(U, M, Z) do { if (small) launch_tiles(h, k_tile_build<U, M, 128, Z>, "k_tile_build", 128, bsm, B, nt); \
                                   else launch_tiles(h, k_tile_build<U, M, 256, Z>, "k_tile_build", 256, bsm, B, nt); } while (0)
      if (mp) { if (!B.uni) BUILD_LAUNCH(false, true, true); else if (zones) BUILD_LAUNCH(true, true, true); else BUILD_LAUNCH(true, true, false); }
      else { if (!B.uni) BUILD_LAUNCH(false, false, true); else if (zones) BUILD_LAUNCH(true, false, true); else BUILD_LAUNCH(true, false, false); }
#undef BUILD_LAUNCH
    }
    CK(cudaMemcpyAsync(h->h_flags, h->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->h_flags + 11, h->d_flags + 11, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    int mx = h->h_flags[0];
    if (mx <= h->stride) { h->maxneigh = std::max<long long>(h->maxneigh, h->h_flags[11]); break; }
    // head room of 40 %: a longer row costs address space, not traffic (rows are read up to their counts), while every overflow later in
    // the run costs a second build plus a cudaFree / cudaMalloc of both row arrays (seen as 10-170 ms outliers of one build, gpurun r02p / r02q)
    if (getenv("B200_VERBOSE")) fprintf(stderr, "b200: step %lld, longest row %d > stride %d: rebuilding with longer rows\n", (long long)h->ntimestep, mx, h->stride);
    h->stride = ((int)(mx * 1.4) + 8 + 31) / 32 * 32;
    if (attempt == 7) throw std::string("b200: neighbor row overflow");
  }
  return true;
}

// ------------------------------------------------------------- reneighbor ---
// Atom::setup_sort_bins (atom.cpp:1659-1730): called by Atom::setup at the start of every run and whenever the box changes
static void setup_sort_bins(b200_sph *h)
{
  Geom &g = h->g;
  double binsize = h->sort_binsize > 0.0 ? h->sort_binsize : 0.5 * h->cutneighmax;
  if (binsize == 0.0) throw std::string("Atom sorting has bin size = 0.0");
  double bininv = 1.0 / binsize;
  SortGeom &sg = h->sortgeom;
  double nb = 1.0;
  for (int d = 0; d < 3; d++) {
    sg.lo[d] = g.sublo[d];
    int n = (int)((g.subhi[d] - g.sublo[d]) * bininv);
    if (g.dim == 2 && d == 2) n = 1;
    if (n == 0) n = 1;
    sg.n[d] = n; sg.inv[d] = n / (g.subhi[d] - g.sublo[d]);
    nb *= n;
  }
  if (nb > 2.0e9) throw std::string("Too many atom sorting bins");
  h->sortgeom_ok = true;
}
// Atom::sort (atom.cpp:1555-1650) as a re-numbering of `orig`; slots [0, nslots), dead ones (alive == 0) skipped
static void atom_sort(b200_sph *h, int nslots, const int *alive)
{
  const int B = 256;
  h->nextsort = (h->ntimestep / h->sortfreq) * h->sortfreq + h->sortfreq;
  if (h->shrink || !h->sortgeom_ok) setup_sort_bins(h);          // "if (domain->box_change) setup_sort_bins()" (:1569)
  const SortGeom &sg = h->sortgeom;
  const int nbins = sg.n[0] * sg.n[1] * sg.n[2];
  if (nbins == 1 || !nslots) return;
  h->sort_cnt.ensure(nbins + 2); h->sort_fill.ensure(nbins + 2);
  h->cellid.ensure(nslots + 1); h->perm.ensure(nslots + 1); h->perm2.ensure(nslots + 1); h->key.ensure(nslots + 1);
  ensure_scan_tmp(h, std::max<size_t>(nslots + 2, nbins + 2));
  CK(cudaMemsetAsync(h->sort_cnt.p, 0, (nbins + 2) * sizeof(int), h->st));
  CK(cudaMemsetAsync(h->sort_fill.p, 0, (nbins + 2) * sizeof(int), h->st));
  LAUNCH(h, k_sort_bin, nblk(nslots, B), B, sg, nslots, h->C().xt.p, alive, h->C().orig.p, h->cellid.p, h->sort_cnt.p, h->key.p);
  scan_exclusive(h, h->sort_cnt.p, nbins, h->scan_tmp.p);
  LAUNCH(h, k_scatter, nblk(nslots, B), B, nslots, h->cellid.p, h->sort_cnt.p, h->sort_fill.p, h->perm.p);
  LAUNCH(h, k_sort_segments, nblk((long long)nbins * 32, B), B, nbins, h->sort_cnt.p, h->perm.p, h->perm2.p, h->key.p);
  LAUNCH(h, k_sort_assign, nblk(nslots, B), B, nslots, h->sort_cnt.p + nbins, h->perm2.p, h->C().orig.p);
}


static void neighbor_build(b200_sph *h, bool do_pbc)
{
  Geom &g = h->g;
  const int B = 256;
  int nl = h->nlocal;
  h->tbegin(T_NEIGH_BIN);
  if (h->shrink) refit_box(h);       // shrink-wrapped dimensions are not periodic, so Domain::pbc (below) and the extent commute
  h->ensure_cap(nl + h->nghost, true);
  // 0. multi-rank: wrap, then migrate atoms that left the sub-domain (verlet.cpp:243-250)
  int nslots = nl;
  const int *alive = nullptr;
  if (h->world > 1) {
    h->alive.ensure(nl + 1);
    if (nl) {
      LAUNCH(h, k_pbc, nblk(nl, B), B, g, nl, h->C().xt.p);
      LAUNCH(h, k_fill_int, nblk(nl, B), B, nl, h->alive.p, 1);
    }
    nslots = comm_exchange(h, nl);
    alive = h->alive.p; do_pbc = false;
  }
  // Atom::sort between exchange and borders: at every setup and on the first rebuild at or after nextsort (verlet.cpp:106,251)
  bool sorted_now = false;
  if (h->sortfreq > 0 && (h->sort_pending || (h->setup_done && h->ntimestep >= h->nextsort))) {
    if (h->world == 1 && do_pbc && nl) LAUNCH(h, k_pbc, nblk(nl, B), B, g, nl, h->C().xt.p);      // the sort bins see wrapped positions (domain->pbc comes first)
    atom_sort(h, nslots, alive);
    sorted_now = true; h->sort_pending = false;
  }
  h->cellid.ensure(nslots + 1); h->perm.ensure(nslots + 1); h->perm2.ensure(nslots + 1); h->key.ensure(nslots + 1);
  ensure_scan_tmp(h, std::max<size_t>(nslots + 2, g.ncells + 2));
  // 1. owned atoms -> cell order (dead slots drop out)
  CK(cudaMemsetAsync(h->cso.p, 0, (g.ncells + 2) * sizeof(int), h->st));
  CK(cudaMemsetAsync(h->cellfill.p, 0, (g.ncells + 2) * sizeof(int), h->st));
  if (nslots) {
    LAUNCH(h, k_owned_cells, nblk(nslots, B), B, g, nslots, h->C().xt.p, h->cellid.p, h->cso.p, do_pbc ? 1 : 0, alive);
    scan_exclusive(h, h->cso.p, g.ncells, h->scan_tmp.p);
    if (h->world > 1) {
      CK(cudaMemcpyAsync(h->h_flags + 4, h->cso.p + g.ncells, sizeof(int), cudaMemcpyDeviceToHost, h->st));
      CK(cudaStreamSynchronize(h->st));
      nl = h->h_flags[4];
    }
    LAUNCH(h, k_scatter, nblk(nslots, B), B, nslots, h->cellid.p, h->cso.p, h->cellfill.p, h->perm.p);
    LAUNCH(h, k_owned_keys, nblk(nslots, B), B, nslots, h->C().tag.p, h->key.p);
    LAUNCH(h, k_sort_segments, nblk((long long)g.ncells * 32, B), B, g.ncells, h->cso.p, h->perm.p, h->perm2.p, h->key.p);
    OwnedSet &a = h->S[h->cur], &b = h->S[h->cur ^ 1];
    b.ensure(std::max(nslots, nl), false, h->st);
    if (nl) LAUNCH(h, k_permute_owned, nblk(nl, B), B, g, nl, h->perm2.p, a.view(), b.view(), h->multiphase, h->key.p);
    h->cur ^= 1;
  }
  h->nlocal = nl;
  if (sorted_now) h->next_orig = nl;                 // the local indices are 0 .. nlocal-1 again
  if (nl) LAUNCH(h, k_fill_int, nblk(nl, B), B, nl, h->gimage.p, 13);
  // 2. ghosts: the staged x, y, z swaps of CommBrick::borders, then a cell-ordered index of them
  comm_borders(h);
  int ng = h->nghost;
  CK(cudaMemsetAsync(h->csg.p, 0, (g.ncells + 2) * sizeof(int), h->st));
  if (ng) {
    h->gcell.ensure(ng); h->gperm.ensure(ng); h->gorder.ensure(ng); h->gkey.ensure(ng);
    CK(cudaMemsetAsync(h->cellfill.p, 0, (g.ncells + 2) * sizeof(int), h->st));
    LAUNCH(h, k_ghost_cells, nblk(ng, B), B, g, nl, ng, h->C().xt.p, h->C().tag.p, h->gimage.p, h->gcell.p, h->csg.p, h->gkey.p);
    scan_exclusive(h, h->csg.p, g.ncells, h->scan_tmp.p);
    LAUNCH(h, k_scatter, nblk(ng, B), B, ng, h->gcell.p, h->csg.p, h->cellfill.p, h->gperm.p);
    LAUNCH(h, k_sort_segments, nblk((long long)g.ncells * 32, B), B, g.ncells, h->csg.p, h->gperm.p, h->gorder.p, h->gkey.p);
  }
  h->ensure_cap(nl + ng, true);
  if ((h->check || h->far_margin > 0.0) && nl + ng) { h->xhold.ensure((size_t)3 * (nl + ng)); LAUNCH(h, k_store_xhold, nblk(nl + ng, B), B, nl + ng, h->C().xt.p, h->xhold.p); }      // ghosts of a rank without owned atoms included (k_unpack_forward reads them)
  h->tend();
  // 3. rows
  h->tbegin(T_NEIGH_BUILD);
  int na = h->nall();
  if (h->tile_on && !tile_rows(h)) { h->tile_on = false; h->tile_ok = false; }
  for (int attempt = 0; attempt < 8 && na && !h->tile_on; attempt++) {
    h->nbr.ensure((size_t)((na + 31) / 32) * 32 * h->stride); h->far.ensure((size_t)((na + 31) / 32) * 32 * h->stride);
    CK(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(int), h->st));
    BuildArgs A;
    A.g = g; A.nlocal = nl; A.nghost = h->nghost; A.stride = h->stride; A.ntypes1 = h->ntypes + 1;
    A.xt = h->C().xt.p; A.orig = h->C().orig.p; A.cso = h->cso.p; A.csg = h->csg.p; A.gorder = h->gorder.p; A.cutneighsq = h->d_cutneighsq.p; A.prunesq = h->d_prunesq.p; A.farsq = h->d_farsq.p; A.far = h->far.p; A.numfar = h->numfar.p;
    A.nbr = h->nbr.p; A.numneigh = h->numneigh.p; A.maxcount = h->d_flags; A.hbn = h->half_bin ? 1 : 0;
    LAUNCH(h, k_build, nblk(g.ncells, BUILD_WARPS), BUILD_WARPS * 32, A);
    CK(cudaMemcpyAsync(h->h_flags, h->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    int mx = h->h_flags[0];
    h->maxneigh = std::max<long long>(h->maxneigh, mx);
    if (mx <= h->stride) break;
    h->stride = ((int)(mx * 1.2) + 8 + 31) / 32 * 32;     // the reference's hard cap is oneatom = 2000 (neighbor.cpp:81)
    if (attempt == 7) throw std::string("b200: neighbor row overflow");
  }
  h->tend();
  h->rows_tiled = h->tile_on;
  if (zones_on(h)) {       // cell of every row, displacement bounds of the cells back to zero, every tile back to its near rows
    h->rowcell.ensure(nl + 1); h->celld.ensure(g.ncells + 2); h->tzone.ensure(h->ntiles + 1);
    if (nl) LAUNCH(h, k_row_cells, nblk(nl, B), B, nl, h->perm2.p, h->cellid.p, h->rowcell.p);
    CK(cudaMemsetAsync(h->celld.p, 0, (g.ncells + 2) * sizeof(unsigned), h->st));
    CK(cudaMemsetAsync(h->tzone.p, 0, h->ntiles + 1, h->st));
  }
  CK(cudaMemsetAsync(h->d_dmaxsq, 0, sizeof(unsigned long long), h->st));
  CK(cudaMemsetAsync(h->d_scan_far, 0, 3 * sizeof(int), h->st));
  h->ago = 0; h->nbuilds++;
}

// ------------------------------------------------------------- pair plan ----
static int kind_of(int style)
{
  switch (style) {
  case B200_PAIR_TAITWATER: return K_TAIT;
  case B200_PAIR_TAITWATER_MORRIS: return K_MORRIS;
  case B200_PAIR_TAITWATER_MULTIPHASE: return K_TAITMP;
  case B200_PAIR_SURFACETENSION: return K_SURF;
  case B200_PAIR_HEATCONDUCTION: return K_HEAT;
  case B200_PAIR_HEATCONDUCTION_MULTIPHASE: return K_HEATMP;
  case B200_PAIR_HEATCONDUCTION_PHASECHANGE: return K_HEATPC;
  case B200_PAIR_IDEALGAS: return K_IDEAL;
  case B200_PAIR_LJ: return K_LJ;
  default: return 0;
  }
}
static const int FUSED[] = {K_TAIT | K_HEAT, K_MORRIS | K_HEAT, K_TAITMP | K_SURF, K_TAITMP | K_SURF | K_HEATMP, K_TAITMP | K_SURF | K_HEATPC,
                            K_TAITMP | K_HEATMP, K_TAITMP | K_HEATPC, K_SURF | K_HEATMP, K_SURF | K_HEATPC};
static bool fusable(int kinds)
{
  for (int f : FUSED) if (f == kinds) return true;
  return (kinds & (kinds - 1)) == 0;
}
static void build_plan(b200_sph *h)
{
  h->plan.clear();
  double psq[MAXTT];
  for (int k = 0; k < MAXTT; k++) { psq[k] = -1.0; for (int s = 0; s < h->npair; s++) psq[k] = std::max(psq[k], h->h_tab[s].cutsq[k]); }
  // far rows: entries at least skin/4 outside the largest pair cutoff (only when there is a skin)
  double fsq[MAXTT];
  h->far_margin = h->skin > 0.0 ? 0.25 * h->skin : 0.0;
  for (int k = 0; k < MAXTT; k++) fsq[k] = (h->far_margin > 0.0 && psq[k] >= 0.0) ? (sqrt(psq[k]) + h->far_margin) * (sqrt(psq[k]) + h->far_margin) : 1e300;
  // tile rows: a mid zone from cut + skin/16 to cut + skin/4 (b200_tile.cuh)
  double msq[MAXTT];
  h->mid_margin = 0.25 * h->far_margin;
  for (int k = 0; k < MAXTT; k++) msq[k] = (h->mid_margin > 0.0 && psq[k] >= 0.0) ? (sqrt(psq[k]) + h->mid_margin) * (sqrt(psq[k]) + h->mid_margin) : 1e300;
  memcpy(h->h_midsq, msq, sizeof msq);
  h->d_midsq.ensure(MAXTT);
  CK(cudaMemcpyAsync(h->d_midsq.p, msq, sizeof msq, cudaMemcpyHostToDevice, h->st));
  h->d_prunesq.ensure(MAXTT); h->d_farsq.ensure(MAXTT);
  memcpy(h->h_farsq, fsq, sizeof fsq);
  CK(cudaMemcpyAsync(h->d_farsq.p, fsq, sizeof fsq, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->d_prunesq.p, psq, sizeof psq, cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  int k = 0;
  while (k < h->npair) {
    int st = h->h_tab[k].style;
    Pass p{}; p.nslots = 1; p.slots[0] = k;
    if (st == B200_PAIR_RHOSUM) { p.type = 0; h->plan.push_back(p); k++; continue; }
    if (st == B200_PAIR_RHOSUM_MULTIPHASE) { p.type = 1; h->plan.push_back(p); k++; continue; }
    if (st == B200_PAIR_COLORGRADIENT) { p.type = 2; h->plan.push_back(p); k++; continue; }
    // greedy group of consecutive force-type sub-styles that has a fused instantiation
    p.type = 3; p.kinds = kind_of(st); k++;
    while (k < h->npair && p.nslots < 3) {
      int kk = kind_of(h->h_tab[k].style);
      if (!kk || (p.kinds & kk) || !fusable(p.kinds | kk) || ((p.kinds | kk) & K_LJ)) break;
      p.kinds |= kk; p.slots[p.nslots++] = k; k++;
    }
    h->plan.push_back(p);
  }
  for (const Pass &p : h->plan)
    if (p.type == 3 && (p.kinds & K_LJ)) {
      bool full = false;
      for (const Pass &q : h->plan) if (q.type != 3) full = true;
      if (!full || h->multiphase)
        throw std::string("pair sph/lj/b200 needs atom_style meso and a full-list sub-style (sph/rhosum) in the deck: without one LAMMPS builds its half "
                          "list in another order (half_bin_newton), and the style's result depends on that order (pair_sph_lj.cpp:139)");
    }
  // Without a full-list sub-style (sph/rhosum, sph/rhosum/multiphase, sph/colorgradient) the reference derives nothing from a full list:
  // its half list is half_bin_newton's (neigh_half_bin.cpp:285-420), with another pair ownership than half_from_full_newton's
  h->half_bin = true;
  for (const Pass &p : h->plan) if (p.type != 3) h->half_bin = false;
  // tile path (b200_tile.cuh): single-phase decks, and multiphase decks without fix phase_change
  bool ok = h->tile_ok && !getenv("B200_NO_TILE") && !h->plan.empty();
  if (h->multiphase && getenv("B200_NO_TILE_MP")) ok = false;
  if (h->multiphase && h->half_bin) ok = false;      // the multiphase tile entries carry half_from_full_newton's ownership; k_build knows both rules
  int np = 2, nk = 1;
  const int SP = K_TAIT | K_MORRIS | K_HEAT | K_IDEAL, MPK = K_TAITMP | K_SURF | K_HEATMP | K_HEATPC;
  for (const Pass &p : h->plan) {
    if (p.type == 0) { if (h->multiphase) ok = false; continue; }
    if (p.type == 1 || p.type == 2) { if (!h->multiphase) ok = false; continue; }
    if (p.type != 3) { ok = false; break; }
    if (!h->multiphase && (p.kinds & K_IDEAL)) {      // asymmetric viscosity table (see k_force): needs the half-list orientation, which the single-phase tile rows do not carry
      const PairTab &T = h->h_tab[p.slots[0]];
      for (int i = 1; i <= h->ntypes; i++) for (int j = 1; j <= h->ntypes; j++) if (T.visc[i * MAXT1 + j] != T.visc[j * MAXT1 + i]) ok = false;
    }
    if (!h->multiphase && !(p.kinds & ~SP)) {
      bool fluid = (p.kinds & (K_TAIT | K_MORRIS | K_IDEAL)) != 0, heat = (p.kinds & K_HEAT) != 0;
      np = std::max(np, fluid ? (heat ? 5 : 4) : 3); nk = std::max(nk, (fluid ? 1 : 0) + (heat ? 1 : 0));
    } else if (h->multiphase && !(p.kinds & ~MPK)) {
      int mask = 0x03 | ((p.kinds & K_TAITMP) ? 0x1c : 0) | ((p.kinds & K_SURF) ? (h->g.dim == 3 ? 0x160 : 0x60) : 0) | ((p.kinds & (K_HEATMP | K_HEATPC)) ? 0x90 : 0);
      np = std::max(np, __builtin_popcount(mask));
      nk = std::max(nk, ((p.kinds & K_TAITMP) ? 1 : 0) + ((p.kinds & K_SURF) ? 1 : 0) + ((p.kinds & (K_HEATMP | K_HEATPC)) ? 1 : 0));
    } else { ok = false; break; }
  }
  h->tile_on = ok; h->tile_nparts = np; h->tile_nk = nk;
  long long capb = (long long)TILE_SMEM_MAX - nk * (long long)sizeof(PairTab) - 2 * (long long)sizeof(TileDesc) - 64;
  h->tile_slotcap = (int)std::min<long long>(std::min<long long>(h->multiphase ? TMP_MAXSLOTS : TILE_MAXSLOTS, capb / (16 * np)), TILE_MAXSLOTS) & ~1;
  if (const char *e = getenv("B200_TILE_SLOTCAP")) h->tile_slotcap = std::max(2, std::min(h->tile_slotcap, atoi(e)) & ~1);   // tests: force small tiles / the row-path fall-back
  if (const char *e = getenv("B200_TILE_SPLIT")) h->tile_split = atoi(e);
  if (h->tile_split != 1 && h->tile_split != 2 && h->tile_split != 4) h->tile_split = 2;    // 4: density pass only
}

static PairArgs pair_args(b200_sph *h)
{
  PairArgs A{};
  OwnedSet &c = h->C();
  A.nlocal = h->nlocal; A.nall = h->nall(); A.stride = h->stride; A.dim = h->g.dim; A.multiphase = h->multiphase; A.nrec = 1;
  A.list = h->nbr.p; A.cnt = h->numneigh.p; A.far = h->far.p; A.numfar = h->numfar.p; A.scan_far = h->d_scan_far;
  A.xt = c.xt.p; A.vr = c.vr.p; A.vm = c.vm.p; A.cgm = c.cgm.p; A.rec = h->rec.p; A.e = c.e.p; A.cv = c.cv.p;
  A.vr_out = c.vr.p; A.cg_out = c.cgm.p; A.fd = c.fd.p; A.de = c.de.p;
  return A;
}

template <int KINDS> static void launch_force(b200_sph *h, PairArgs &A)
{
  int grid = nblk(A.nall, PAIR_THREADS);
  if (h->g.dim == 3) k_force<KINDS, true><<<grid, PAIR_THREADS, 0, h->st>>>(A);
  else k_force<KINDS, false><<<grid, PAIR_THREADS, 0, h->st>>>(A);
  post_launch(h, "k_force");
}

static TileArgs tile_args(b200_sph *h, int pstride)
{
  TileArgs A{};
  OwnedSet &c = h->C();
  A.nlocal = h->nlocal; A.ngrp = h->stride / 8; A.pstride = pstride; A.cap = h->tile_cap;
  A.rec = h->trec.p; A.near = (const uint4 *)h->nbr.p; A.far = (const uint4 *)h->far.p; A.numneigh = h->numneigh.p; A.numfar = h->numfar.p;
  A.scan_far = h->d_scan_far; A.tiles = h->tiles.p; A.ntiles = h->ntiles; A.counter = h->d_tflags + 4;
  A.xt = c.xt.p; A.vr_out = c.vr.p; A.fd = c.fd.p; A.de = c.de.p;
  A.vm = c.vm.p; A.cg_out = c.cgm.p; A.gorder = h->gorder.p; A.dim = h->g.dim;
  A.virow = h->virow.p;
  return A;
}
// which: 0 all records, 1 owned atoms only, 2 ghosts only (halo overlap: the ghosts' records wait for the halo)
static int tile_records(b200_sph *h, int nparts, int force, int epart, const PairTab *fluid, int which = 0)
{
  int na = h->nall(), pstride = (na + 7) & ~7;
  h->trec.ensure((size_t)pstride * nparts);
  OwnedSet &c = h->C();
  int i0 = which == 2 ? h->nlocal : 0, i1 = which == 1 ? h->nlocal : na;
  TileRecArgs R{h->nlocal, na, pstride, force, epart, i0, i1, h->gorder.p, c.xt.p, c.vr.p, c.e.p, fluid, h->trec.p};
  if (i1 > i0) LAUNCH(h, k_tile_records, nblk(i1 - i0, 256), 256, R);
  return pstride;
}
// ---- halo overlap: the halo runs on a second stream between two events; the pair pass that follows evaluates the interior
//      tiles first and only then waits for it (far_flags: the ghosts' displacement enters the zone flags after the wait)
static void far_flags(b200_sph *h);
static void halo_wait(b200_sph *h)
{
  if (!h->comm_pending) return;
  CK(cudaStreamWaitEvent(h->st, h->ev_comm, 0));
  h->comm_pending = false;
  far_flags(h);
}
template <class F> static void halo_async(b200_sph *h, F comm)
{
  CK(cudaEventRecord(h->ev_main, h->st));
  CK(cudaStreamWaitEvent(h->st2, h->ev_main, 0));
  std::swap(h->st, h->st2);                 // every launch / NCCL call of `comm` goes to the halo stream
  try { comm(); } catch (...) { std::swap(h->st, h->st2); throw; }
  std::swap(h->st, h->st2);
  CK(cudaEventRecord(h->ev_comm, h->st2));
  h->comm_pending = true;
}
static bool overlap_ok(const b200_sph *h) { return h->rows_tiled && !h->multiphase && h->nint > 0 && !h->no_overlap && (h->nghost || h->world > 1); }
// one pair pass over the tiles: all at once, or interior tiles | wait for the halo | ghost records | boundary tiles
template <class Launch, class Rec> static void tile_pass(b200_sph *h, TileArgs &A, Launch launch, Rec records)
{
  const unsigned char *tz = zones_on(h) ? h->tzone.p : nullptr;
  A.tzone = tz;
  if (!h->comm_pending) { A.pstride = records(0); A.rec = h->trec.p; A.tiles = h->tiles.p; A.ntiles = h->ntiles; launch(A, 0); return; }
  A.pstride = records(1); A.rec = h->trec.p;
  A.tiles = h->tiles.p; A.ntiles = h->nint;
  launch(A, h->world > 1 ? 8 : 2);           // a few SMs stay free for the halo's kernels (NCCL send/recv, pack, unpack)
  halo_wait(h);
  records(2);
  A.tiles = h->tiles.p + h->nint; A.ntiles = h->ntiles - h->nint; A.tzone = tz ? tz + h->nint : nullptr;
  launch(A, 0);
}
// constants of a sub-style whose coefficients do not depend on the type pair (TileUni); false if they do
static bool tile_uni(const b200_sph *h, const PairTab &T, TileUni &U)
{
  memset(&U, 0, sizeof U);
  bool first = true, ok = true;
  unsigned used = 0;
  for (int i = 1; i <= h->ntypes; i++)
    for (int j = 1; j <= h->ntypes; j++) {
      int k = i * MAXT1 + j;
      if (T.cutsq[k] < 0.0) continue;
      U.mapmask |= 1ull << k; used |= (1u << i) | (1u << j);
      if (first) { U.cutsq = T.cutsq[k]; U.h = T.h[k]; U.c0 = T.c0[k]; U.c1 = T.c1[k]; U.visc = T.visc[k]; first = false; }
      else if (U.cutsq != T.cutsq[k] || U.h != T.h[k] || U.c0 != T.c0[k] || U.c1 != T.c1[k] || U.visc != T.visc[k]) ok = false;
    }
  if (first) return false;
  bool f2 = true;
  for (int i = 1; i <= h->ntypes; i++) {
    if (!(used & (1u << i))) continue;
    if (f2) { U.mass = T.mass[i]; U.cs = T.cs[i]; U.self = T.self0[i]; f2 = false; }
    else if (U.mass != T.mass[i] || U.cs != T.cs[i] || U.self != T.self0[i]) ok = false;
  }
  return ok;
}
template <int KINDS> static void launch_tile_force(b200_sph *h, const TileArgs &A, bool uni, int reserve)
{
  constexpr bool F = (KINDS & (K_TAIT | K_MORRIS | K_IDEAL)) != 0, H = (KINDS & K_HEAT) != 0;
  constexpr int NP = F ? (H ? 5 : 4) : 3, NK = (F ? 1 : 0) + (H ? 1 : 0);
  size_t smem = TileSmem<NP, NK>::bytes(h->tile_cap);
  if (h->vir_now && F) {            // thermo step: the instantiation that also sums the rows' virial
    if (uni) launch_tiles(h, k_tile_force<KINDS, 2, true, true>, "k_tile_force", TILE_ROWS * 2, smem, A, A.ntiles, reserve);
    else launch_tiles(h, k_tile_force<KINDS, 2, false, true>, "k_tile_force", TILE_ROWS * 2, smem, A, A.ntiles, reserve);
  } else if (h->tile_split == 1) {
    if (uni) launch_tiles(h, k_tile_force<KINDS, 1, true, false>, "k_tile_force", TILE_ROWS, smem, A, A.ntiles, reserve);
    else launch_tiles(h, k_tile_force<KINDS, 1, false, false>, "k_tile_force", TILE_ROWS, smem, A, A.ntiles, reserve);
  } else {
    if (uni) launch_tiles(h, k_tile_force<KINDS, 2, true, false>, "k_tile_force", TILE_ROWS * 2, smem, A, A.ntiles, reserve);
    else launch_tiles(h, k_tile_force<KINDS, 2, false, false>, "k_tile_force", TILE_ROWS * 2, smem, A, A.ntiles, reserve);
  }
}
static int tile_records_mp(b200_sph *h, int mode, const PairTab *fluid)
{
  int na = h->nall(), pstride = (na + 7) & ~7;
  h->trec.ensure((size_t)pstride * (mode ? TILE_MP_NPART : 2));
  OwnedSet &c = h->C();
  TileRecMpArgs R{h->nlocal, na, pstride, mode, h->g.dim, h->gorder.p, c.xt.p, c.vr.p, c.cgm.p, c.e.p, c.cv.p, fluid, h->trec.p};
  LAUNCH(h, k_tile_records_mp, nblk(na, 256), 256, R);
  return pstride;
}
// cutoff / kernel constants of a sub-style that are the same for all its mapped type pairs (viscosity, alpha, D may still differ)
static bool tile_uni_geo(const b200_sph *h, const PairTab &T, TileUni &U)
{
  memset(&U, 0, sizeof U);
  bool first = true, ok = true;
  for (int i = 1; i <= h->ntypes; i++)
    for (int j = 1; j <= h->ntypes; j++) {
      int k = i * MAXT1 + j;
      if (T.cutsq[k] < 0.0) continue;
      U.mapmask |= 1ull << k;
      if (first) { U.cutsq = T.cutsq[k]; U.h = T.h[k]; U.c0 = T.c0[k]; U.c1 = T.c1[k]; first = false; }
      else if (U.cutsq != T.cutsq[k] || U.h != T.h[k] || U.c0 != T.c0[k] || U.c1 != T.c1[k]) ok = false;
    }
  return ok && !first;
}
template <int KINDS> static void launch_tile_force_mp(b200_sph *h, TileArgs &A, bool gu)
{
  const bool d3 = h->g.dim == 3 || !(KINDS & K_SURF);
  size_t smem = d3 ? TileSmem<MpParts<KINDS, true>::n, MpParts<KINDS, true>::nk>::bytes(h->tile_cap) : TileSmem<MpParts<KINDS, false>::n, MpParts<KINDS, false>::nk>::bytes(h->tile_cap);
  {                                            // owned rows, then the ghost rows, in one tile queue (tile_rows)
    const int nt = h->ntiles + h->ngtiles;
    A.ntiles = nt;
    if (d3) {
      if (gu) launch_tiles(h, k_tile_force_mp<KINDS, true, true>, "k_tile_force_mp", TILE_MP_NT, smem, A, nt);
      else launch_tiles(h, k_tile_force_mp<KINDS, true, false>, "k_tile_force_mp", TILE_MP_NT, smem, A, nt);
    } else {
      if (gu) launch_tiles(h, k_tile_force_mp<KINDS, false, true>, "k_tile_force_mp", TILE_MP_NT, smem, A, nt);
      else launch_tiles(h, k_tile_force_mp<KINDS, false, false>, "k_tile_force_mp", TILE_MP_NT, smem, A, nt);
    }
  }
}
static void run_pass_tile_mp(b200_sph *h, const Pass &p)
{
  halo_wait(h);
  if (p.type == 1 || p.type == 2) {
    const PairTab &T = h->h_tab[p.slots[0]];
    bool active = T.nstep != 0 && (h->ntimestep % T.nstep) == 0;
    if (!active) return;
    h->tbegin(p.type == 1 ? T_DENSITY : T_COLORGRAD);
    int pstride = p.type == 1 ? tile_records(h, 2, 0, -1, nullptr) : tile_records_mp(h, 0, nullptr);
    TileArgs A = tile_args(h, pstride);
    A.tab[0] = h->d_tab[p.slots[0]];
    const bool gu = tile_uni_geo(h, T, A.uni[0]) && !h->tile_nouni;
    size_t smem = TileSmem<2, 1>::bytes(h->tile_cap);
    if (p.type == 1) {
      if (gu) launch_tiles(h, k_tile_full_mp<0, true>, "k_tile_rhosum_mp", TILE_MPFULL_NT, smem, A);
      else launch_tiles(h, k_tile_full_mp<0, false>, "k_tile_rhosum_mp", TILE_MPFULL_NT, smem, A);
    } else {
      if (gu) launch_tiles(h, k_tile_full_mp<1, true>, "k_tile_colorgradient", TILE_MPFULL_NT, smem, A);
      else launch_tiles(h, k_tile_full_mp<1, false>, "k_tile_colorgradient", TILE_MPFULL_NT, smem, A);
    }
    h->tend();
    return;
  }
  // force pass: canonical table order fluid, surf, heat
  int nk = 0; const PairTab *fluid = nullptr;
  TileArgs A = tile_args(h, 0);
  const int wants[3] = {K_TAITMP, K_SURF, K_HEATMP | K_HEATPC};
  bool gu = !h->tile_nouni;
  for (int want : wants)
    for (int s = 0; s < p.nslots; s++)
      if (kind_of(h->h_tab[p.slots[s]].style) & want) {
        gu = tile_uni_geo(h, h->h_tab[p.slots[s]], A.uni[nk]) && gu;
        if (nk && A.uni[nk].c1 != A.uni[0].c1) gu = false;      // one kernel-derivative evaluation serves all sub-styles only if they share h
        A.tab[nk++] = h->d_tab[p.slots[s]]; if (want & K_TAITMP) fluid = h->d_tab[p.slots[s]];
      }
  h->tbegin(T_DERIVE);
  A.pstride = tile_records_mp(h, 1, fluid);
  A.rec = h->trec.p;
  h->tend();
  h->tbegin(T_FORCE);
  A.accum = h->f_clean ? 0 : 1; h->f_clean = false;
  switch (p.kinds) {
  case K_TAITMP: launch_tile_force_mp<K_TAITMP>(h, A, gu); break;
  case K_SURF: launch_tile_force_mp<K_SURF>(h, A, gu); break;
  case K_HEATMP: launch_tile_force_mp<K_HEATMP>(h, A, gu); break;
  case K_HEATPC: launch_tile_force_mp<K_HEATPC>(h, A, gu); break;
  case K_TAITMP | K_SURF: launch_tile_force_mp<K_TAITMP | K_SURF>(h, A, gu); break;
  case K_TAITMP | K_SURF | K_HEATMP: launch_tile_force_mp<K_TAITMP | K_SURF | K_HEATMP>(h, A, gu); break;
  case K_TAITMP | K_SURF | K_HEATPC: launch_tile_force_mp<K_TAITMP | K_SURF | K_HEATPC>(h, A, gu); break;
  case K_TAITMP | K_HEATMP: launch_tile_force_mp<K_TAITMP | K_HEATMP>(h, A, gu); break;
  case K_TAITMP | K_HEATPC: launch_tile_force_mp<K_TAITMP | K_HEATPC>(h, A, gu); break;
  case K_SURF | K_HEATMP: launch_tile_force_mp<K_SURF | K_HEATMP>(h, A, gu); break;
  case K_SURF | K_HEATPC: launch_tile_force_mp<K_SURF | K_HEATPC>(h, A, gu); break;
  default: throw std::string("b200: no tile force kernel for this multiphase sub-style group");
  }
  h->tend();
}

static void run_pass_tile(b200_sph *h, const Pass &p)
{
  const int B = 256;
  if (h->multiphase) { run_pass_tile_mp(h, p); return; }
  if (p.type == 0) {
    const PairTab &T = h->h_tab[p.slots[0]];
    bool active = T.nstep != 0 && (h->ntimestep % T.nstep) == 0;     // pair_sph_rhosum.cpp:112-113
    h->tbegin(T_DENSITY);
    if (active) {
      TileArgs A = tile_args(h, 0);
      A.tab[0] = h->d_tab[p.slots[0]];
      bool uni = tile_uni(h, T, A.uni[0]) && !h->tile_nouni;
      size_t smem = uni ? TileSmem<2, 0>::bytes(h->tile_cap) : TileSmem<2, 1>::bytes(h->tile_cap);
      // lanes per row: 2 when two 512-thread CTAs then share an SM (one evaluates while the other sits at its tile barrier / bulk copies:
      // C2 density 0.256 -> 0.211 ms, gpurun r02p), else 4 lanes in one 1024-thread CTA (round 1: beats 2 lanes in one CTA per SM)
      int dsplit = 4;
      if (uni) {
        int occ = 0;
        if (smem <= tile_optin((const void *)k_tile_rhosum<2, true>)) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tile_rhosum<2, true>, TILE_ROWS * 2, smem));
        if (occ >= 2) dsplit = 2;
      }
      if (getenv("B200_TILE_SPLIT")) dsplit = h->tile_split;
      if (const char *e = getenv("B200_DENSITY_SPLIT")) dsplit = atoi(e);
      tile_pass(h, A,
        [&](TileArgs &a, int reserve) {
          if (dsplit == 1) {
            if (uni) launch_tiles(h, k_tile_rhosum<1, true>, "k_tile_rhosum", TILE_ROWS, smem, a, a.ntiles, reserve);
            else launch_tiles(h, k_tile_rhosum<1, false>, "k_tile_rhosum", TILE_ROWS, smem, a, a.ntiles, reserve);
          } else if (dsplit == 2) {
            if (uni) launch_tiles(h, k_tile_rhosum<2, true>, "k_tile_rhosum", TILE_ROWS * 2, smem, a, a.ntiles, reserve);
            else launch_tiles(h, k_tile_rhosum<2, false>, "k_tile_rhosum", TILE_ROWS * 2, smem, a, a.ntiles, reserve);
          } else {
            if (uni) launch_tiles(h, k_tile_rhosum<4, true>, "k_tile_rhosum", TILE_ROWS * 4, smem, a, a.ntiles, reserve);
            else launch_tiles(h, k_tile_rhosum<4, false>, "k_tile_rhosum", TILE_ROWS * 4, smem, a, a.ntiles, reserve);
          }
        },
        [&](int which) { return tile_records(h, 2, 0, -1, nullptr, which); });
    } else halo_wait(h);
    if (h->nghost || h->world > 1) {      // comm->forward_comm_pair (:203): the ghosts' new rho (decided from global state: peers may wait for this rank)
      auto rho_halo = [&]() {
        comm_forward_generic(h, 1,
          [&](Swap &s, double *dst) { LAUNCH(h, k_pack_rho, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, h->C().vr.p, dst); },
          [&](Swap &s, double *buf) { LAUNCH(h, k_unpack_rho, nblk(s.nrecv, B), B, s.nrecv, s.firstrecv, h->C().vr.p, buf); });
      };
      if (overlap_ok(h)) halo_async(h, rho_halo); else rho_halo();
    }
    h->tend();
    return;
  }
  const PairTab *fluid = nullptr, *heat = nullptr, *hfluid = nullptr, *hheat = nullptr;
  for (int s = 0; s < p.nslots; s++) {
    int kk = kind_of(h->h_tab[p.slots[s]].style);
    if (kk & (K_TAIT | K_MORRIS | K_IDEAL)) { fluid = h->d_tab[p.slots[s]]; hfluid = &h->h_tab[p.slots[s]]; }
    if (kk & K_HEAT) { heat = h->d_tab[p.slots[s]]; hheat = &h->h_tab[p.slots[s]]; }
  }
  int nparts = fluid ? (heat ? 5 : 4) : 3;
  TileArgs A = tile_args(h, 0);
  int nk = 0;
  bool uni = !h->tile_nouni;
  if (fluid) { uni = tile_uni(h, *hfluid, A.uni[nk]) && uni; A.tab[nk++] = fluid; }
  if (heat) { uni = tile_uni(h, *hheat, A.uni[nk]) && uni; A.tab[nk++] = heat; }
  if (fluid && heat && (A.uni[0].mass != A.uni[1].mass)) uni = false;
  h->tbegin(T_FORCE);
  A.accum = h->f_clean ? 0 : 1; h->f_clean = false;
  tile_pass(h, A,
    [&](TileArgs &a, int reserve) {
      switch (p.kinds) {
      case K_TAIT: launch_tile_force<K_TAIT>(h, a, uni, reserve); break;
      case K_MORRIS: launch_tile_force<K_MORRIS>(h, a, uni, reserve); break;
      case K_HEAT: launch_tile_force<K_HEAT>(h, a, uni, reserve); break;
      case K_IDEAL: launch_tile_force<K_IDEAL>(h, a, uni, reserve); break;
      case K_TAIT | K_HEAT: launch_tile_force<K_TAIT | K_HEAT>(h, a, uni, reserve); break;
      case K_MORRIS | K_HEAT: launch_tile_force<K_MORRIS | K_HEAT>(h, a, uni, reserve); break;
      default: throw std::string("b200: no tile force kernel for this sub-style group");
      }
    },
    [&](int which) { return tile_records(h, nparts, fluid ? 1 : 0, heat ? (fluid ? 4 : 2) : -1, fluid, which); });
  h->tend();
}

static void run_pass(b200_sph *h, const Pass &p)
{
  // a rank without owned atoms still takes part in every halo of the pass (it may hold ghosts and send lists);
  // only its local kernels have nothing to do (launch_tiles / the row kernels return on empty ranges)
  if (h->tile_on != h->rows_tiled) throw std::string("b200: the neighbor rows were built for another pair plan (call b200_setup / b200_reneighbor)");
  if (h->tile_on) { run_pass_tile(h, p); return; }
  halo_wait(h);
  const int B = 256;
  if (p.type <= 2) {
    const PairTab &T = h->h_tab[p.slots[0]];
    bool active = T.nstep != 0 && (h->ntimestep % T.nstep) == 0;     // pair_sph_rhosum.cpp:112-113
    PairArgs A = pair_args(h);
    A.tab[0] = h->d_tab[p.slots[0]];
    int grid = nblk(A.nlocal, PAIR_THREADS);
    if (p.type == 0) {
      h->tbegin(T_DENSITY);
      if (active) LAUNCH(h, k_rhosum<false>, grid, PAIR_THREADS, A);
      if (h->nghost || h->world > 1)       // comm->forward_comm_pair (:203)
        comm_forward_generic(h, 1,
          [&](Swap &s, double *dst) { LAUNCH(h, k_pack_rho, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, h->C().vr.p, dst); },
          [&](Swap &s, double *buf) { LAUNCH(h, k_unpack_rho, nblk(s.nrecv, B), B, s.nrecv, s.firstrecv, h->C().vr.p, buf); });
      h->tend();
    } else if (p.type == 1) {
      h->tbegin(T_DENSITY);
      if (active) LAUNCH(h, k_rhosum<true>, grid, PAIR_THREADS, A);
      h->tend();
    } else if (active) {
      h->tbegin(T_DERIVE);
      LAUNCH(h, k_records, nblk(A.nall, B), B, A.nall, 2, 1, 0, (const PairTab *)nullptr, A.xt, A.vr, A.cgm, A.e, A.cv, h->rec.p);
      h->tend();
      h->tbegin(T_COLORGRAD);
      LAUNCH(h, k_colorgradient, grid, PAIR_THREADS, A);
      h->tend();
    }
    return;
  }
  if (p.kinds == K_LJ) {          // sph/lj: its own kernel (owner-side evaluation in the reference's list order, b200_lj.cuh)
    OwnedSet &c = h->C();
    h->tbegin(T_FORCE);
    h->f_clean = false;
    CK(cudaMemsetAsync(h->d_flags + 13, 0, sizeof(int), h->st));
    LjArgs L{h->nlocal, h->stride, h->g.dim, h->g.sx, h->g.sy, h->g.sz, h->nbr.p, h->far.p, h->numneigh.p, h->numfar.p,
             c.xt.p, c.vr.p, c.e.p, c.cv.p, c.orig.p, h->gimage.p, c.fd.p, c.de.p, h->d_tab[p.slots[0]], h->d_flags + 13};
    if (h->nlocal) LAUNCH(h, k_force_lj, nblk(h->nlocal, 128), 128, L);
    CK(cudaMemcpyAsync(h->h_flags + 13, h->d_flags + 13, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (h->h_flags[13]) throw std::string("pair sph/lj/b200: more than 512 in-cutoff half-list neighbors of one atom");
    h->tend();
    return;
  }
  // force pass: canonical table order fluid, surf, heat
  PairArgs A = pair_args(h);
  int nk = 0; const PairTab *fluid = nullptr;
  const int wants[3] = {K_TAIT | K_MORRIS | K_TAITMP | K_IDEAL, K_SURF, K_HEAT | K_HEATMP | K_HEATPC};
  for (int want : wants)
    for (int s = 0; s < p.nslots; s++)
      if (kind_of(h->h_tab[p.slots[s]].style) & want) { A.tab[nk++] = h->d_tab[p.slots[s]]; if (want & K_TAIT) fluid = h->d_tab[p.slots[s]]; }
  bool mp = (p.kinds & (K_TAITMP | K_SURF | K_HEATMP | K_HEATPC)) != 0;
  int nrec = mp ? 4 : (((p.kinds & K_HEAT) && fluid) ? 3 : 2);
  h->tbegin(T_DERIVE);
  LAUNCH(h, k_records, nblk(A.nall, B), B, A.nall, mp ? 1 : 0, nrec, (!mp && !fluid) ? 1 : 0, fluid, A.xt, A.vr, A.cgm, A.e, A.cv, h->rec.p);
  h->tend();
  A.nrec = nrec;
  h->tbegin(T_FORCE);
  h->f_clean = false;
  switch (p.kinds) {
  case K_TAIT: launch_force<K_TAIT>(h, A); break;
  case K_MORRIS: launch_force<K_MORRIS>(h, A); break;
  case K_HEAT: launch_force<K_HEAT>(h, A); break;
  case K_TAITMP: launch_force<K_TAITMP>(h, A); break;
  case K_SURF: launch_force<K_SURF>(h, A); break;
  case K_HEATMP: launch_force<K_HEATMP>(h, A); break;
  case K_HEATPC: launch_force<K_HEATPC>(h, A); break;
  case K_IDEAL: launch_force<K_IDEAL>(h, A); break;
  case K_TAIT | K_HEAT: launch_force<K_TAIT | K_HEAT>(h, A); break;
  case K_MORRIS | K_HEAT: launch_force<K_MORRIS | K_HEAT>(h, A); break;
  case K_TAITMP | K_SURF: launch_force<K_TAITMP | K_SURF>(h, A); break;
  case K_TAITMP | K_SURF | K_HEATMP: launch_force<K_TAITMP | K_SURF | K_HEATMP>(h, A); break;
  case K_TAITMP | K_SURF | K_HEATPC: launch_force<K_TAITMP | K_SURF | K_HEATPC>(h, A); break;
  case K_TAITMP | K_HEATMP: launch_force<K_TAITMP | K_HEATMP>(h, A); break;
  case K_TAITMP | K_HEATPC: launch_force<K_TAITMP | K_HEATPC>(h, A); break;
  case K_SURF | K_HEATMP: launch_force<K_SURF | K_HEATMP>(h, A); break;
  case K_SURF | K_HEATPC: launch_force<K_SURF | K_HEATPC>(h, A); break;
  default: throw std::string("b200: no force kernel instantiation for this sub-style group");
  }
  h->tend();
}

// ---------------------------------------------------------- Verlet stages ---
static void force_clear(b200_sph *h)
{
  int na = h->nall();
  if (!na) return;
  CK(cudaMemsetAsync(h->C().fd.p, 0, (size_t)na * sizeof(double4), h->st));
  CK(cudaMemsetAsync(h->C().de.p, 0, (size_t)na * sizeof(double), h->st));
  h->f_clean = true;
}
// Pair::virial_fdotr_compute (pair.cpp:1403-1451) of the pair forces just computed, before the reverse halo
static void pair_compute_all(b200_sph *h)
{
  const bool rowsum = h->vir_now && h->rows_tiled && !h->multiphase;      // single-phase tile path: no forces on ghosts, the rows sum their pairs (k_tile_force VIR)
  if (rowsum && h->nlocal) { h->virow.ensure((size_t)h->nlocal * 6); CK(cudaMemsetAsync(h->virow.p, 0, (size_t)h->nlocal * 6 * sizeof(double), h->st)); }
  for (const Pass &p : h->plan) run_pass(h, p);
  if (!h->vir_now) return;
  halo_wait(h);
  h->virpart.ensure(VIR_BLOCKS * 6 + 8);
  if (rowsum) LAUNCH(h, k_virial_partial, VIR_BLOCKS, VIR_THREADS, h->nlocal, (const double4 *)nullptr, (const double4 *)nullptr, (const double *)h->virow.p, h->virpart.p);
  else LAUNCH(h, k_virial_partial, VIR_BLOCKS, VIR_THREADS, h->nall(), (const double4 *)h->C().xt.p, (const double4 *)h->C().fd.p, (const double *)nullptr, h->virpart.p);
  LAUNCH(h, k_virial_final, 1, 32, h->virpart.p, h->virpart.p + VIR_BLOCKS * 6);
  CK(cudaMemcpyAsync(h->h_vir, h->virpart.p + VIR_BLOCKS * 6, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  h->vir_now = false;
}
static void post_final(b200_sph *h, int rev, int post, int fin)
{
  const int B = 256;
  halo_wait(h);
  h->f_clean = false;
  h->tbegin(T_FINAL);
  if (rev && (h->nghost || h->world > 1) && (!h->tile_on || h->multiphase))      // the single-phase tile path puts nothing on ghosts (b200_tile.cuh)
    comm_reverse_generic(h, NB_REVERSE,
      [&](Swap &s, double *dst) { LAUNCH(h, k_pack_reverse, nblk(s.nrecv, B), B, s.nrecv, s.firstrecv, h->comm_arrays(), dst); },
      [&](Swap &s, double *buf) { LAUNCH(h, k_unpack_reverse, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, h->comm_arrays(), buf); });
  if ((post || fin) && h->nlocal)
    LAUNCH(h, k_post_final, nblk(h->nlocal, B), B, h->nlocal, h->fl, h->step_arrays(), 0.5 * h->dt * h->ftm2v, post, fin,
           h->dtreset ? h->d_dt : (const double *)nullptr, 0.5 * h->ftm2v, h->d_progs.p, (double)h->ntimestep, h->dt);
  h->tend();
}
static void initial_integrate(b200_sph *h)
{
  CK(cudaMemsetAsync(h->d_flags + 1, 0, sizeof(int), h->st));      // the moved flag feeds an all-reduce (neigh_decide): reset it on empty ranks too
  if (!h->nlocal) return;
  h->tbegin(T_INTEGRATE);
  int track = h->far_margin > 0.0;
  LAUNCH(h, k_initial_integrate, nblk(h->nlocal, 256), 256, h->nlocal, h->fl, h->step_arrays(), h->dt, 0.5 * h->dt * h->ftm2v, h->check,
         h->xhold.p, h->triggersq, h->d_flags + 1, track, h->d_dmaxsq, h->dtreset ? h->d_dt : (const double *)nullptr, 0.5 * h->ftm2v,
         h->rowcell.p, zones_on(h) ? h->celld.p : (unsigned *)nullptr);
  h->tend();
}
static void forward_comm(b200_sph *h)
{
  if (!h->nghost && h->world == 1) return;
  const int B = 256;
  h->tbegin(T_COMM);
  comm_forward_generic(h, fwd_width(h->multiphase, h->ghost_velocity),
    [&](Swap &s, double *dst) { LAUNCH(h, k_pack_forward, nblk(s.nsend, B), B, s.nsend, s.sendlist.p, h->comm_arrays(), s.dim, s.shift, dst, h->multiphase, h->ghost_velocity); },
    [&](Swap &s, double *buf) { LAUNCH(h, k_unpack_forward, nblk(s.nrecv, B), B, s.nrecv, s.firstrecv, h->comm_arrays(), buf, h->multiphase, h->ghost_velocity,
                                       h->far_margin > 0.0 ? h->xhold.p : (const double *)nullptr, h->d_dmaxsq, h->gcell.p,
                                       zones_on(h) ? h->celld.p : (unsigned *)nullptr, h->nlocal); });
  h->tend();
}
// far / mid rows must be scanned once 2 * dmax reaches their margin; dmax = largest displacement since the build over the owned atoms
// (k_initial_integrate) and the ghosts (k_unpack_forward): every candidate of a row is one of the two, so the bound is rank-local
static void far_flags(b200_sph *h)
{
  if (h->far_margin > 0.0 && h->nlocal)
    LAUNCH(h, k_far_flag, 1, 1, h->d_dmaxsq, h->far_margin * h->far_margin, h->mid_margin * h->mid_margin, h->d_scan_far);
  if (zones_on(h))
    LAUNCH(h, k_tile_zone, nblk((long long)h->ntiles * 32, 128), 128, h->tiles.p, h->ntiles, h->celld.p, h->far_margin * h->far_margin, h->mid_margin * h->mid_margin, h->tzone.p);
}
// Neighbor::decide, neighbor.cpp:1332-1347 (+ check_distance :1360-1410)
static int neigh_decide(b200_sph *h)
{
  for (const PcFix &f : h->pcs) if (h->ntimestep == f.next) return 1;   // fix->next_reneighbor (neighbor.cpp:1334-1338)
  h->ago++;
  if (h->ago >= h->delay && h->ago % h->every == 0) {
    if (!h->check) return 1;
    if (h->world > 1) NCK(g_nccl.AllReduce(h->d_flags + 1, h->d_flags + 1, 1, ncclInt, ncclMax, h->nccl, h->st));   // MPI_Allreduce MAX (neighbor.cpp:1407)
    CK(cudaMemcpyAsync(h->h_flags + 1, h->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    int flag = h->h_flags[1];
    if (flag && h->ago == std::max(h->every, h->delay)) h->ndanger++;
    return flag;
  }
  return 0;
}
static void dt_download(b200_sph *h);
// FixPhaseChange::pre_exchange (fix_phase_change.cpp:167-352); runs on the rows of the previous build
static void phase_change(b200_sph *h, PcFix &f)
{
  if (f.next != h->ntimestep) return;
  f.next += f.d.nfreq;
  int nl = h->nlocal, na = h->nall();
  if (!nl && h->world == 1) return;
  dt_download(h);
  h->tbegin(T_PHASE);
  // a rank without owned atoms has no candidates, but its ghosts' (zero) dmass still travels back and it
  // takes part in the tag_extend gather below: only the local kernels are skipped
  int nins = 0;
  h->pc_dmass.ensure(na + 1);
  OwnedSet &c = h->C();
  PcArrays a{c.xt.p, c.vr.p, c.vm.p, c.cgm.p, c.e.p, c.cv.p, c.orig.p, h->nbr.p, h->far.p, h->numneigh.p, h->numfar.p,
             h->rows_tiled ? 1 : 0, h->stride / 8, h->tiles.p, h->rowtile.p, h->gorder.p};
  if (nl) {
    int norig = std::max(h->next_orig, nl);
    h->pc_flag.ensure(norig); h->pc_thr.ensure(norig); h->pc_dev.ensure(norig); h->pc_new.ensure(PC_MAXNEW);
    CK(cudaMemsetAsync(h->pc_flag.p, 0, norig, h->st));
    PcParams P; P.d = f.d; P.dim = h->g.dim; P.nlocal = nl; P.nall = na; P.norig = norig; P.stride = h->stride; P.dt = h->dt;
    for (int d = 0; d < 3; d++) { P.sublo[d] = h->g.sublo[d]; P.subhi[d] = h->g.subhi[d]; P.boxhi[d] = h->g.boxhi[d]; }
    LAUNCH(h, k_pc_candidates, nblk(na, 128), 128, P, a, h->pc_flag.p, h->pc_thr.p, h->pc_dev.p, h->pc_dmass.p);
    CK(cudaMemsetAsync(f.d_state + 1, 0, 2 * sizeof(int), h->st));
    // compact the candidates (ascending local index) so the serial walk touches only them
    h->pos.ensure(norig + 2); h->flag.ensure(norig + 2); ensure_scan_tmp(h, norig + 2);
    LAUNCH(h, k_pc_mark, nblk(norig, 256), 256, norig, h->pc_flag.p, h->pos.p);
    scan_exclusive(h, h->pos.p, norig, h->scan_tmp.p);
    CK(cudaMemcpyAsync(h->h_flags + 7, h->pos.p + norig, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    int ncand = h->h_flags[7];
    if (ncand) LAUNCH(h, k_pc_compact, nblk(norig, 256), 256, norig, h->pc_flag.p, h->pos.p, h->flag.p);
    LAUNCH(h, k_pc_walk, 1, 32, P, a, h->pc_flag.p, h->pc_thr.p, h->pc_dev.p, h->pc_dmass.p, h->pc_new.p, f.d_state, h->flag.p, ncand);
    CK(cudaMemcpyAsync(h->h_flags + 5, f.d_state + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  } else if (na) CK(cudaMemsetAsync(h->pc_dmass.p, 0, (size_t)na * sizeof(double), h->st));
  if (h->nghost || h->world > 1) comm_reverse_scalar_add(h, h->pc_dmass.p);      // comm->reverse_comm_fix (:324)
  if (nl) {
    LAUNCH(h, k_pc_apply, nblk(nl, 256), 256, nl, a, h->pc_dmass.p);
    CK(cudaStreamSynchronize(h->st));
    nins = h->h_flags[5];
    if (h->h_flags[6]) throw std::string("fix phase_change: more than PC_MAXNEW insertions in one call");
  }
  int tag0 = h->maxtag, ninsall = nins;
  if (h->world > 1) {        // Atom::tag_extend: MPI_Scan of the per-rank counts (atom.cpp:598-630)
    std::vector<int> all(h->world);
    h->h_flags[12] = nins;
    CK(cudaMemcpyAsync(h->d_flags + 12, h->h_flags + 12, sizeof(int), cudaMemcpyHostToDevice, h->st));
    DevBuf<int> tmp; tmp.ensure(h->world);
    NCK(g_nccl.AllGather(h->d_flags + 12, tmp.p, 1, ncclInt, h->nccl, h->st));
    CK(cudaMemcpyAsync(all.data(), tmp.p, h->world * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    tmp.release();
    ninsall = 0;
    for (int r = 0; r < h->world; r++) { if (r < h->rank) tag0 += all[r]; ninsall += all[r]; }
  }
  if (nins > 0) {
    h->ensure_cap((size_t)nl + nins, true);
    OwnedSet &cc = h->C();
    AppendArrays ap{cc.xt.p, cc.vr.p, cc.vm.p, cc.fd.p, cc.cgm.p, cc.e.p, cc.de.p, cc.cv.p, cc.tag.p, cc.mask.p, cc.orig.p};
    LAUNCH(h, k_pc_append, nblk(nins, 128), 128, nl, nins, h->pc_new.p, ap, f.d.to_type, f.d.to_mass, f.d.groupbit, tag0, h->next_orig);
    h->nlocal += nins; h->next_orig += nins; h->ninserted += nins;
  }
  if (ninsall > 0) { h->nghost = 0; h->maxtag += ninsall; }
  h->tend();
}
// FixDtReset::end_of_step
static void dt_reset(b200_sph *h)
{
  const unsigned long long big = 0x7ff0000000000000ull;     // +inf: larger than any timestep
  CK(cudaMemcpyAsync((unsigned long long *)h->d_dt + 1, &big, sizeof big, cudaMemcpyHostToDevice, h->st));
  if (h->nlocal) LAUNCH(h, k_dt_min, nblk(h->nlocal, 256), 256, h->nlocal, h->dtr_bit, h->C().mask.p, h->C().vm.p, h->C().fd.p, h->dtr_xmax, h->ftm2v,
                        (unsigned long long *)h->d_dt + 1);
  if (h->world > 1) NCK(g_nccl.AllReduce((unsigned long long *)h->d_dt + 1, (unsigned long long *)h->d_dt + 1, 1, ncclUint64, ncclMin, h->nccl, h->st));   // MPI_Allreduce MIN (:171)
  LAUNCH(h, k_dt_apply, 1, 1, (const unsigned long long *)h->d_dt + 1, h->dtr_minbound, h->dtr_tmin, h->dtr_maxbound, h->dtr_tmax, h->d_dt, h->ntimestep);
}
static void dt_download(b200_sph *h)
{
  if (!h->dtreset) return;
  CK(cudaMemcpyAsync(h->h_dtv, h->d_dt, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  h->dt = h->h_dtv[0]; h->atime = h->h_dtv[2];
  memcpy(&h->atimestep, h->h_dtv + 3, sizeof(long long)); memcpy(&h->laststep, h->h_dtv + 4, sizeof(long long));
}
static void reneighbor(b200_sph *h)
{
  for (PcFix &f : h->pcs) phase_change(h, f);        // modify->pre_exchange (verlet.cpp:241)
  neighbor_build(h, true);
}

static void do_setup(b200_sph *h)
{
  if (!h->geom_ready) setup_geometry(h);
  if (h->world > 1) {      // new-atom tags continue after the GLOBAL maximum (Atom::tag_extend, atom.cpp:603-605)
    h->h_flags[12] = h->maxtag;
    CK(cudaMemcpyAsync(h->d_flags + 12, h->h_flags + 12, sizeof(int), cudaMemcpyHostToDevice, h->st));
    NCK(g_nccl.AllReduce(h->d_flags + 12, h->d_flags + 12, 1, ncclInt, ncclMax, h->nccl, h->st));
    CK(cudaMemcpyAsync(h->h_flags + 12, h->d_flags + 12, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->maxtag = h->h_flags[12];
  }
  build_plan(h);
  // The single-phase tile path evaluates a pair with a ghost on both owners' sides, which equals the reference's one evaluation as long
  // as the ghost's fields are fresh copies of its owner's.  Through the setup force evaluation they are not when setup_pre_force changes
  // vest (k_vest_stale): the pair must then be evaluated where the reference's half list holds it, ghost rows and reverse halo included,
  // which is what the row path does.  The setup evaluation of such a run is taken on the row path; the rows of the tile path are built
  // right after it (same atoms, same ghosts, now with their owners' vest -- which the reference's ghosts receive with the forward halo
  // of step 1 before anything reads them).  Shipped deck: examples/USER/sph/cavity_flow; fixtures cavity2d, cavity2d_rhosum.
  bool row_setup = false;
  if (h->tile_on && !h->multiphase && (h->world > 1 || h->g.periodic[0] || h->g.periodic[1] || h->g.periodic[2])) {
    CK(cudaMemsetAsync(h->d_flags + 14, 0, sizeof(int), h->st));
    if (h->nlocal) LAUNCH(h, k_vest_stale, nblk(h->nlocal, 256), 256, h->nlocal, h->fl, h->step_arrays(), h->d_flags + 14);
    if (h->world > 1) NCK(g_nccl.AllReduce(h->d_flags + 14, h->d_flags + 14, 1, ncclInt, ncclMax, h->nccl, h->st));
    CK(cudaMemcpyAsync(h->h_flags + 14, h->d_flags + 14, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (h->h_flags[14]) { h->tile_on = false; row_setup = true; }
  }
  if (h->sortfreq > 0) { h->sortgeom_ok = false; h->sort_pending = true; }      // Atom::setup -> setup_sort_bins; Verlet::setup: if (atom->sortfreq > 0) atom->sort()
  neighbor_build(h, true);
  h->nbuilds = 0;
  h->vir_now = h->vir_request; h->vir_request = false;
  force_clear(h);
  if (h->nlocal) LAUNCH(h, k_setup_pre_force, nblk(h->nlocal, 256), 256, h->nlocal, h->fl, h->step_arrays());
  // ghosts carry vest of the border comm (before setup_pre_force), exactly as in Verlet::setup
  pair_compute_all(h);
  post_final(h, 1, 1, 0);
  if (row_setup && !getenv("B200_STALE_SETUP_STAYS_ON_ROWS")) {      // (the switch keeps the whole run on the row path: A/B in tests/test_gpu_tile.py)
    h->tile_on = true;
    neighbor_build(h, true);       // falls back to rows by itself (collectively) if a tile does not fit
    h->nbuilds = 0;
  }
  if (h->dtreset) {            // FixDtReset::setup -> end_of_step
    h->h_dtv[0] = h->dt; h->h_dtv[1] = 0.0; h->h_dtv[2] = h->atime;
    memcpy(h->h_dtv + 3, &h->atimestep, sizeof(long long)); memcpy(h->h_dtv + 4, &h->laststep, sizeof(long long));
    CK(cudaMemcpyAsync(h->d_dt, h->h_dtv, 5 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));                   // h_dtv is read back into by dt_download right after
    dt_reset(h); dt_download(h);
  }
  h->setup_done = true;
}
static void do_run(b200_sph *h, int n)
{
  if (!h->setup_done) throw std::string("b200_run called before b200_setup");
  for (int s = 0; s < n; s++) {
    h->ntimestep++;
    initial_integrate(h);
    if (neigh_decide(h)) reneighbor(h);
    else if (overlap_ok(h)) { far_flags(h); halo_async(h, [&]() { forward_comm(h); }); }     // interior tiles start on the owned atoms' displacement bound
    else { forward_comm(h); far_flags(h); }
    if (s == n - 1 && h->vir_request) { h->vir_now = true; h->vir_request = false; }
    force_clear(h);
    pair_compute_all(h);
    post_final(h, 1, 1, 1);
    if (h->dtreset && h->ntimestep % h->dtr_every == 0) dt_reset(h);      // modify->end_of_step
    h->nsteps++;
  }
  dt_download(h);
  if (h->timing) h->tflush();
}

// ------------------------------------------------------------ table fill ----
static void fill_tab(b200_sph *h, const b200_pair_desc *d, PairTab &T)
{
  memset(&T, 0, sizeof T);
  int n1 = h->ntypes + 1, dim = h->g.dim;
  const double n3 = 0.0716197243913529, n2 = 0.04195297663091802;   // sph_kernel_quintic.cpp
  const double nq = dim == 3 ? n3 : n2;
  T.style = d->style; T.nstep = d->nstep; T.kind = kind_of(d->style);
  for (int k = 0; k < MAXTT; k++) T.cutsq[k] = -1.0;
  for (int t = 0; t < MAXT1; t++) { T.iskip[t] = 1; T.mass[t] = h->mass[t]; }
  bool guni = true; double g0 = 0; bool gset = false;
  for (int i = 1; i < n1; i++) {
    if (d->rho0) T.rho0[i] = d->rho0[i];
    if (d->B) T.B[i] = d->B[i];
    if (d->soundspeed) T.cs[i] = d->soundspeed[i];
    if (d->gamma) { T.gamma[i] = d->gamma[i]; }
    if (d->rbackground) T.rb[i] = d->rbackground[i];
    for (int j = 1; j < n1; j++) {
      int s = i * n1 + j, k = i * MAXT1 + j;
      if (!d->mapped[s]) continue;
      T.iskip[i] = 0;
      double hh = d->cut[s], ih = 1.0 / hh, ihsq = ih * ih;
      T.cutsq[k] = d->cutsq[s]; T.h[k] = hh;
      switch (d->style) {
      case B200_PAIR_RHOSUM:
        T.c1[k] = ihsq;
        T.c0[k] = dim == 3 ? 2.1541870227086614782e0 * ihsq * ih : 1.5915494309189533576e0 * ihsq;
        break;
      case B200_PAIR_RHOSUM_MULTIPHASE:
        T.c1[k] = ih; T.c0[k] = dim == 3 ? nq * ih * ih * ih : nq * ih * ih;
        break;
      case B200_PAIR_COLORGRADIENT: case B200_PAIR_SURFACETENSION:
        T.c1[k] = ih; T.c0[k] = dim == 3 ? 3.0 * nq * ih * ih * ih * ih : 3.0 * nq * ih * ih * ih;
        if (d->alpha) T.visc[k] = d->alpha[s];
        break;
      case B200_PAIR_TAITWATER: case B200_PAIR_TAITWATER_MORRIS: case B200_PAIR_HEATCONDUCTION: case B200_PAIR_IDEALGAS: case B200_PAIR_LJ:
        T.c0[k] = dim == 3 ? -25.066903536973515383e0 * ihsq * ihsq * ihsq * ih : -19.098593171027440292e0 * ihsq * ihsq * ihsq;
        T.visc[k] = d->style == B200_PAIR_HEATCONDUCTION ? (d->alpha ? d->alpha[s] : 0.0) : (d->viscosity ? d->viscosity[s] : 0.0);
        break;
      case B200_PAIR_TAITWATER_MULTIPHASE: case B200_PAIR_HEATCONDUCTION_MULTIPHASE: case B200_PAIR_HEATCONDUCTION_PHASECHANGE:
        T.c1[k] = ih; T.c0[k] = dim == 3 ? 3.0 * nq * ih * ih * ih * ih : 3.0 * nq * ih * ih * ih;
        T.visc[k] = d->style == B200_PAIR_TAITWATER_MULTIPHASE ? (d->viscosity ? d->viscosity[s] : 0.0) : (d->alpha ? d->alpha[s] : 0.0);
        if (d->tc) T.tc[k] = d->tc[s];
        if (d->fixflag) T.fixflag[k] = d->fixflag[s];
        break;
      default: throw std::string("b200_pair_add: unknown pair style");
      }
    }
    // self terms use h = cut[itype][itype]
    int sii = i * n1 + i;
    if (d->mapped[sii]) {
      double hh = d->cut[sii];
      if (d->style == B200_PAIR_RHOSUM) T.self0[i] = dim == 3 ? 2.1541870227086614782 / (hh * hh * hh) : 1.5915494309189533576e0 / (hh * hh);
      if (d->style == B200_PAIR_RHOSUM_MULTIPHASE) T.self0[i] = dim == 3 ? (nq * 66.0) / (hh * hh * hh) : (nq * 66.0) / (hh * hh);
    }
    if (d->gamma && !T.iskip[i]) { if (!gset) { g0 = d->gamma[i]; gset = true; } else if (d->gamma[i] != g0) guni = false; }
  }
  T.gamma_uniform = guni ? 1 : 0;
}

// =================================================================== ABI ====
extern "C" {

const char *b200_last_error(void) { return g_err.c_str(); }
const char *b200_version(void) { return "b200sph 0.1 (sm_100a)"; }

int b200_create(b200_sph **out, int device)
{
  API_BEGIN
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) throw std::string("b200_create: no CUDA device (the SPH hot path has no CPU fallback)");
  if (device < 0 || device >= n) throw std::string("b200_create: bad device index");
  CK(cudaSetDevice(device));
  b200_sph *h = new b200_sph();
  h->device = device;
  CK(cudaMalloc(&h->d_flags, 16 * sizeof(int)));
  CK(cudaMemset(h->d_flags, 0, 16 * sizeof(int)));
  CK(cudaMalloc(&h->d_dmaxsq, sizeof(unsigned long long))); CK(cudaMemset(h->d_dmaxsq, 0, sizeof(unsigned long long)));
  h->d_scan_far = h->d_flags + 8;
  CK(cudaMalloc(&h->d_tflags, 16 * sizeof(int))); CK(cudaMemset(h->d_tflags, 0, 16 * sizeof(int)));
  CK(cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  h->nsm = prop.multiProcessorCount;
  CK(cudaMalloc(&h->d_red, 24 * sizeof(double))); CK(cudaMallocHost(&h->h_red, 16 * sizeof(double)));
  CK(cudaMallocHost(&h->h_flags, 32 * sizeof(int)));
  CK(cudaMallocHost(&h->h_vir, 8 * sizeof(double))); memset(h->h_vir, 0, 8 * sizeof(double));
  memset(&h->fl, 0, sizeof h->fl);
  *out = h;
  API_END
}
int b200_destroy(b200_sph *h)
{
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->S[0].release(); h->S[1].release(); h->pc_flag.release(); h->pc_thr.release(); h->pc_dmass.release(); h->pc_dev.release(); h->pc_new.release(); for (PcFix &f : h->pcs) cudaFree(f.d_state); for (PcFix &f : h->pcs_old) cudaFree(f.d_state);
  h->rec.release(); h->far.release(); h->numfar.release(); h->d_prunesq.release(); h->d_farsq.release(); h->d_midsq.release(); cudaFree(h->d_dmaxsq); h->gimage.release();
  h->cellid.release(); h->perm.release(); h->perm2.release(); h->gcell.release(); h->gperm.release(); h->gorder.release(); h->flag.release(); h->pos.release(); h->alive.release();
  h->sendbuf.release(); h->recvbuf.release(); for (int q = 0; q < 2; q++) { h->xs[q].release(); h->xr[q].release(); } for (int k = 0; k < 6; k++) h->swaps[k].sendlist.release();
  if (h->nccl) g_nccl.CommDestroy(h->nccl);
  cudaFree(h->d_red); cudaFreeHost(h->h_red); if (h->d_dt) cudaFree(h->d_dt); if (h->h_dtv) cudaFreeHost(h->h_dtv);
  h->key.release(); h->gkey.release(); h->cso.release(); h->csg.release();
  h->cellfill.release(); h->scan_tmp.release(); h->xhold.release(); h->stage_d.release(); h->stage_i.release(); h->d_mass.release(); h->nbr.release(); h->numneigh.release(); h->d_cutneighsq.release();
  for (int k = 0; k < MAXPAIR; k++) if (h->d_tab[k]) cudaFree(h->d_tab[k]);
  for (auto &p : h->ev_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  cudaFree(h->d_flags); cudaFreeHost(h->h_flags); cudaFreeHost(h->h_vir); h->virow.release(); h->virpart.release();
  if (h->st2) cudaStreamDestroy(h->st2); if (h->ev_main) cudaEventDestroy(h->ev_main); if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  h->celld.release(); h->rowcell.release(); h->tzone.release(); h->sort_cnt.release(); h->sort_fill.release();
  h->tiles.release(); h->gtiles.release(); h->rowtile.release(); h->trec.release(); cudaFree(h->d_tflags);
  delete h;
  return 0;
}

int b200_comm_unique_id(char id[128])
{
  API_BEGIN
  g_nccl.load();
  static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId size");
  ncclUniqueId u;
  NCK(g_nccl.GetUniqueId(&u));
  memset(id, 0, 128); memcpy(id, &u, sizeof u);
  API_END
}
int b200_comm_init(b200_sph *h, int world, int rank, const int procgrid[3], const int myloc[3], const int procneigh[6], const char id[128])
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  if (world < 1 || rank < 0 || rank >= world || procgrid[0] * procgrid[1] * procgrid[2] != world) throw std::string("b200_comm_init: bad rank / processor grid");
  h->world = world; h->rank = rank;
  for (int d = 0; d < 3; d++) { h->procgrid[d] = procgrid[d]; h->myloc[d] = myloc[d]; h->procneigh[d][0] = procneigh[2 * d]; h->procneigh[d][1] = procneigh[2 * d + 1]; }
  if (world > 1) {
    g_nccl.load();
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    NCK(g_nccl.CommInitRank(&h->nccl, world, u, rank));
  }
  h->geom_ready = false;
  API_END
}

int b200_domain(b200_sph *h, int dim, const double boxlo[3], const double boxhi[3], const int periodicity[3], const double sublo[3],
                const double subhi[3])
{
  API_BEGIN
  if (dim != 2 && dim != 3) throw std::string("b200_domain: dimension must be 2 or 3");
  h->g.dim = dim;
  for (int d = 0; d < 3; d++) {
    h->g.boxlo[d] = boxlo[d]; h->g.boxhi[d] = boxhi[d]; h->g.periodic[d] = periodicity[d];
    h->g.sublo[d] = sublo ? sublo[d] : boxlo[d]; h->g.subhi[d] = subhi ? subhi[d] : boxhi[d];
    if ((h->g.sublo[d] != boxlo[d] || h->g.subhi[d] != boxhi[d]) && h->world == 1) throw std::string("b200_domain: a sub-domain smaller than the box needs b200_comm_init first");
  }
  h->have_domain = true; h->geom_ready = false;
  API_END
}
int b200_boundary(b200_sph *h, const int boundary[6], const double small[3], const double minbox[6])
{
  API_BEGIN
  if (!h->have_domain) throw std::string("b200_boundary: call b200_domain first");
  h->shrink = false;
  for (int d = 0; d < 3; d++) {
    for (int k = 0; k < 2; k++) {
      int b = boundary[2 * d + k];
      if (b < 0 || b > 3) throw std::string("b200_boundary: style must be 0 (p), 1 (f), 2 (s) or 3 (m)");
      if ((b == 0) != (h->g.periodic[d] != 0)) throw std::string("b200_boundary: styles disagree with the periodicity given to b200_domain");
      h->boundary[d][k] = b; h->minbox[d][k] = minbox ? minbox[2 * d + k] : 0.0;
      if (b >= 2) h->shrink = true;
    }
    h->small[d] = small ? small[d] : 0.0;
  }
  h->geom_ready = false;
  API_END
}
int b200_get_box(b200_sph *h, double boxlo[3], double boxhi[3])
{
  API_BEGIN
  for (int d = 0; d < 3; d++) { boxlo[d] = h->g.boxlo[d]; boxhi[d] = h->g.boxhi[d]; }
  API_END
}
int b200_atom_style(b200_sph *h, int multiphase, int ntypes, const double *mass)
{
  API_BEGIN
  if (ntypes < 1 || ntypes >= MAXT1) throw std::string("b200_atom_style: 1..7 atom types supported");
  h->multiphase = multiphase; h->ntypes = ntypes;
  for (int t = 0; t < MAXT1; t++) h->mass[t] = (mass && t <= ntypes) ? mass[t] : 0.0;
  API_END
}
int b200_neighbor(b200_sph *h, double skin, int every, int delay, int check, const double *cutneighsq, double cutneighmax, double cutghost)
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  int n1 = h->ntypes + 1;
  h->skin = skin; h->every = every; h->delay = delay; h->check = check;
  memset(h->h_cutneighsq, 0, sizeof h->h_cutneighsq);
  for (int i = 1; i < n1; i++) for (int j = 1; j < n1; j++) h->h_cutneighsq[i * MAXT1 + j] = cutneighsq[i * n1 + j];
  h->d_cutneighsq.ensure(MAXTT);
  CK(cudaMemcpy(h->d_cutneighsq.p, h->h_cutneighsq, sizeof h->h_cutneighsq, cudaMemcpyHostToDevice));
  h->cutneighmax = cutneighmax; h->g.cutghost = cutghost;
  h->triggersq = 0.25 * skin * skin;      // neighbor.cpp:240
  h->have_neigh = true; h->geom_ready = false;
  API_END
}
int b200_timestep(b200_sph *h, double dt, double ftm2v, long long ntimestep) { h->dt = dt; h->ftm2v = ftm2v; h->ntimestep = ntimestep; return 0; }
int b200_comm_modify(b200_sph *h, int ghost_velocity) { h->ghost_velocity = ghost_velocity; return 0; }
int b200_atom_modify(b200_sph *h, int sortfreq, double userbinsize)
{
  if (sortfreq < 0 || userbinsize < 0.0) return fail("Illegal atom_modify command");
  h->sortfreq = sortfreq; h->sort_binsize = userbinsize; h->sortgeom_ok = false;
  return 0;
}

int b200_pair_clear(b200_sph *h) { h->npair = 0; h->plan.clear(); h->tile_on = false; return 0; }
int b200_pair_add(b200_sph *h, const b200_pair_desc *d)
{
  int slot = -1;
  try {
    CK(cudaSetDevice(h->device));
    if (h->npair == MAXPAIR) throw std::string("b200_pair_add: too many sub-styles");
    if (!h->have_domain || !h->ntypes) throw std::string("b200_pair_add: call b200_domain and b200_atom_style first");
    if (!d->mapped || !d->cut || !d->cutsq) throw std::string("b200_pair_add: mapped, cut and cutsq are required");
    slot = h->npair;
    fill_tab(h, d, h->h_tab[slot]);
    if (!h->d_tab[slot]) CK(cudaMalloc(&h->d_tab[slot], sizeof(PairTab)));
    CK(cudaMemcpy(h->d_tab[slot], &h->h_tab[slot], sizeof(PairTab), cudaMemcpyHostToDevice));
    h->npair++;
  } catch (const std::string &m) { return fail(m); }
  return slot;
}

int b200_fix_clear(b200_sph *h)
{
  memset(&h->fl, 0, sizeof h->fl);
  h->progs.clear();
  // FixPhaseChange keeps next_reneighbor and its RanPark stream across `run` commands (fix_phase_change.cpp:116,345): when the
  // caller re-registers the same fix (VerletB200::configure runs per `run`), b200_fix_phase_change adopts that state again
  for (PcFix &f : h->pcs_old) cudaFree(f.d_state);
  h->pcs_old = h->pcs;
  h->pcs.clear();
  h->dtreset = false;
  return 0;
}
static int add_fix(b200_sph *h, int kind, int bit, double ax, double ay, double az)
{
  if (h->fl.n == MAXFIX) return fail("too many fixes");
  int k = h->fl.n++;
  h->fl.kind[k] = kind; h->fl.bit[k] = bit; h->fl.acc[k][0] = ax; h->fl.acc[k][1] = ay; h->fl.acc[k][2] = az;
  return 0;
}
int b200_fix_meso(b200_sph *h, int groupbit) { return add_fix(h, 1, groupbit, 0, 0, 0); }
int b200_fix_meso_stationary(b200_sph *h, int groupbit) { return add_fix(h, 2, groupbit, 0, 0, 0); }
int b200_fix_gravity(b200_sph *h, int groupbit, double xacc, double yacc, double zacc) { return add_fix(h, 3, groupbit, xacc, yacc, zacc); }
int b200_fix_setmeso(b200_sph *h, int groupbit, int which, double value, int region_kind, const double region[6], int match_inside)
{
  if (which < 0 || which > 2 || region_kind < 0 || region_kind > 2) return fail("b200_fix_setmeso: bad arguments");
  if (add_fix(h, 4, groupbit, 0, 0, 0)) return -1;
  int k = h->fl.n - 1;
  h->fl.ipar[k][0] = which; h->fl.ipar[k][1] = region_kind; h->fl.ipar[k][2] = match_inside;
  h->fl.par[k][0] = value;
  for (int q = 0; q < 6; q++) h->fl.par[k][1 + q] = (region_kind && region) ? region[q] : 0.0;
  return 0;
}
// a variable formula of the fix being registered -> its program slot (1-based; 0 = none)
static int add_formula(b200_sph *h, const char *text, int *slot)
{
  ExprProg P;
  std::string err = expr_compile(text, P);
  if (!err.empty()) return fail(err + " (formula: " + text + ")");
  h->progs.push_back(P);
  h->d_progs.ensure(h->progs.size());
  CK(cudaMemcpy(h->d_progs.p, h->progs.data(), h->progs.size() * sizeof(ExprProg), cudaMemcpyHostToDevice));
  *slot = (int)h->progs.size();
  return 0;
}
int b200_formula_check(const char *formula, const double atom[12], int type, int id, double step, double dt, double *value)
{
  ExprProg P;
  std::string err = expr_compile(formula, P);
  if (!err.empty()) return fail(err + " (formula: " + formula + ")");
  if (value && atom) {
    ExprIn in{atom[0], atom[1], atom[2], atom[3], atom[4], atom[5], atom[6], atom[7], atom[8], atom[9], step, dt, 0.0, type, id};
    *value = expr_eval(P, in);
  }
  return 0;
}
int b200_fix_setmeso_var(b200_sph *h, int groupbit, int which, const char *formula, int region_kind, const double region[6], int match_inside)
{
  API_BEGIN
  if (!formula) return fail("b200_fix_setmeso_var: no formula");
  if (b200_fix_setmeso(h, groupbit, which, 0.0, region_kind, region, match_inside)) return -1;
  if (add_formula(h, formula, &h->fl.prog[h->fl.n - 1][0])) { h->fl.n--; return -1; }
  API_END
}
int b200_fix_addforce(b200_sph *h, int groupbit, const double value[3], const char *const formula[3])
{
  API_BEGIN
  if (add_fix(h, 8, groupbit, value ? value[0] : 0.0, value ? value[1] : 0.0, value ? value[2] : 0.0)) return -1;
  const int k = h->fl.n - 1;
  for (int d = 0; d < 3; d++)
    if (formula && formula[d] && add_formula(h, formula[d], &h->fl.prog[k][d])) { h->fl.n--; return -1; }
  API_END
}
int b200_fix_enforce2d(b200_sph *h, int groupbit) { return add_fix(h, 5, groupbit, 0, 0, 0); }
int b200_fix_dt_reset(b200_sph *h, int groupbit, int nevery, int minbound, double tmin, int maxbound, double tmax, double xmax)
{
  API_BEGIN
  if (nevery <= 0 || xmax <= 0.0 || (minbound && tmin < 0.0) || (maxbound && tmax < 0.0) || (minbound && maxbound && tmin >= tmax))
    throw std::string("Illegal fix dt/reset command");
  CK(cudaSetDevice(h->device));
  if (!h->d_dt) { CK(cudaMalloc(&h->d_dt, 8 * sizeof(double))); CK(cudaMallocHost(&h->h_dtv, 8 * sizeof(double))); }
  h->dtreset = true; h->dtr_bit = groupbit; h->dtr_every = nevery; h->dtr_minbound = minbound; h->dtr_tmin = tmin;
  h->dtr_maxbound = maxbound; h->dtr_tmax = tmax; h->dtr_xmax = xmax;
  API_END
}
int b200_get_timestep(b200_sph *h, double *dt) { *dt = h->dt; return 0; }
int b200_set_time(b200_sph *h, double atime, long long atimestep, long long laststep) { h->atime = atime; h->atimestep = atimestep; h->laststep = laststep; return 0; }
int b200_get_time(b200_sph *h, double *atime, long long *atimestep, long long *laststep)
{ if (atime) *atime = h->atime; if (atimestep) *atimestep = h->atimestep; if (laststep) *laststep = h->laststep; return 0; }
int b200_request_virial(b200_sph *h) { h->vir_request = true; return 0; }
int b200_get_virial(b200_sph *h, double v[6])
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st));
  for (int k = 0; k < 6; k++) v[k] = h->h_vir[k];
  API_END
}
int b200_fix_setmesode(b200_sph *h, int groupbit, double value, int region_kind, const double region[6])
{
  if (region_kind < 0 || region_kind > 2) return fail("b200_fix_setmesode: bad arguments");
  if (add_fix(h, 7, groupbit, 0, 0, 0)) return -1;
  int k = h->fl.n - 1;
  h->fl.ipar[k][1] = region_kind; h->fl.par[k][0] = value;
  for (int q = 0; q < 6; q++) h->fl.par[k][1 + q] = (region_kind && region) ? region[q] : 0.0;
  return 0;
}
int b200_fix_setforce(b200_sph *h, int groupbit, const int set[3], const double value[3])
{
  if (add_fix(h, 6, groupbit, 0, 0, 0)) return -1;
  int k = h->fl.n - 1;
  for (int d = 0; d < 3; d++) { h->fl.ipar[k][d] = set[d] != 0; h->fl.par[k][d] = value[d]; }
  return 0;
}
int b200_fix_phase_change(b200_sph *h, const b200_phase_change_desc *d)
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  if (d->seed <= 0) throw std::string("Illegal value for seed");                   // fix_phase_change.cpp:70
  if (!h->multiphase) throw std::string("fix phase_change requires atom_style meso/multiphase");
  PcFix f; f.d = *d; f.next = d->first_step; f.d_state = nullptr;
  auto same = [](b200_phase_change_desc a, b200_phase_change_desc b) {
    return a.groupbit == b.groupbit && a.Tc == b.Tc && a.Tt == b.Tt && a.Hwv == b.Hwv && a.dr == b.dr && a.to_mass == b.to_mass && a.cutoff == b.cutoff &&
           a.from_type == b.from_type && a.to_type == b.to_type && a.nfreq == b.nfreq && a.seed == b.seed && a.energy_chance_flag == b.energy_chance_flag &&
           a.change_chance == b.change_chance && a.phase_change_rate == b.phase_change_rate && a.maxattempt == b.maxattempt && a.first_step == b.first_step;
  };
  for (size_t k = 0; k < h->pcs_old.size(); k++)
    if (same(h->pcs_old[k].d, *d)) {          // the same fix registered again: keep its next step and RNG position
      f.next = h->pcs_old[k].next; f.d_state = h->pcs_old[k].d_state;
      h->pcs_old.erase(h->pcs_old.begin() + k);
      break;
    }
  if (!f.d_state) {
    CK(cudaMalloc(&f.d_state, 4 * sizeof(int)));
    int st[4] = {d->seed, 0, 0, 0};
    CK(cudaMemcpy(f.d_state, st, sizeof st, cudaMemcpyHostToDevice));
  }
  h->pcs.push_back(f);
  API_END
}

// staging: the caller's AoS arrays are copied verbatim (DMA from pinned memory when the caller
// pinned them) and converted to/from the packed device records by a kernel
static HostMirror stage_layout(b200_sph *h, int n, const b200_atoms *a, size_t *nd_out, size_t *ni_out)
{
  // offsets first (so the buffers can be sized), then pointers
  const size_t NONE = (size_t)-1;
  size_t od = 0, oi = 0;
  auto take = [&](const void *p, size_t &o, size_t w) { if (!p) return NONE; size_t r = o; o += w * (size_t)n; return r; };
  size_t ox = take(a->x, od, 3), ov = take(a->v, od, 3), ove = take(a->vest, od, 3), of = take(a->f, od, 3), oc = take(a->colorgradient, od, 3);
  size_t orho = take(a->rho, od, 1), odrho = take(a->drho, od, 1), oe = take(a->e, od, 1), ode = take(a->de, od, 1), ocv = take(a->cv, od, 1),
         orm = take(a->rmass, od, 1);
  size_t oty = take(a->type, oi, 1), oma = take(a->mask, oi, 1), ota = take(a->tag, oi, 1);
  h->stage_d.ensure(od + 1); h->stage_i.ensure(oi + 1);
  double *bd = h->stage_d.p; int *bi = h->stage_i.p;
  auto D = [&](size_t o) { return o == NONE ? (double *)nullptr : bd + o; };
  auto I = [&](size_t o) { return o == NONE ? (int *)nullptr : bi + o; };
  HostMirror m{D(ox), D(ov), D(ove), D(of), D(oc), D(orho), D(odrho), D(oe), D(ode), D(ocv), D(orm), I(oty), I(oma), I(ota)};
  *nd_out = od; *ni_out = oi;
  return m;
}
static PackArrays pack_arrays(b200_sph *h) { OwnedSet &c = h->C(); return PackArrays{c.xt.p, c.vr.p, c.vm.p, c.fd.p, c.cgm.p, c.e.p, c.de.p, c.cv.p, c.tag.p, c.mask.p, c.orig.p}; }
#define FOR_FIELDS(X) X(x, x, 3) X(v, v, 3) X(vest, vest, 3) X(f, f, 3) X(cg, colorgradient, 3) X(rho, rho, 1) X(drho, drho, 1) X(e, e, 1) X(de, de, 1) X(cv, cv, 1) X(rmass, rmass, 1)

int b200_set_atoms(b200_sph *h, int n, const b200_atoms *a)
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  if (!a->x || !a->type) throw std::string("b200_set_atoms: x and type are required");
  if (n > (int)NBR_INDEX_MASK / 2) throw std::string("b200_set_atoms: too many atoms for 30-bit neighbor indices");
  h->nlocal = n; h->nghost = 0; h->cur = 0;
  h->ensure_cap(n, false);
  h->setup_done = false;
  h->maxtag = n; h->next_orig = n;
  if (!n) return 0;
  size_t nd, ni;
  HostMirror m = stage_layout(h, n, a, &nd, &ni);
#define UP(dev, host, w) if (a->host) CK(cudaMemcpyAsync(m.dev, a->host, (size_t)(w) * n * sizeof(double), cudaMemcpyHostToDevice, h->st));
  FOR_FIELDS(UP)
#undef UP
  if (a->type) CK(cudaMemcpyAsync(m.type, a->type, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->st));
  if (a->mask) CK(cudaMemcpyAsync(m.mask, a->mask, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->st));
  if (a->tag) CK(cudaMemcpyAsync(m.tag, a->tag, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->st));
  h->d_mass.ensure(MAXT1);
  CK(cudaMemcpyAsync(h->d_mass.p, h->mass, sizeof h->mass, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemsetAsync(h->d_flags + 2, 0, 2 * sizeof(int), h->st));       // [2] bad type, [3] largest tag (no host pass over the tags)
  LAUNCH(h, k_pack_atoms, nblk(n, 256), 256, n, m, pack_arrays(h), h->multiphase, h->d_mass.p, h->ntypes, h->d_flags + 2);
  CK(cudaMemcpyAsync(h->h_flags + 2, h->d_flags + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (h->h_flags[2]) throw std::string("b200_set_atoms: atom type out of range");
  h->maxtag = std::max(h->maxtag, h->h_flags[3]);
  API_END
}
int b200_get_natoms(b200_sph *h, int *nlocal, int *nghost) { if (nlocal) *nlocal = h->nlocal; if (nghost) *nghost = h->nghost; return 0; }

// output index of every owned device slot: the LAMMPS local index (one rank), or the rank of that index among the owned atoms
// when migration has made the local-index sequence sparse (several ranks).  On the device: mark the indices in use, scan, look up.
__global__ void k_mark_orig(int n, const int *orig, int *flag) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) flag[orig[i]] = 1; }
__global__ void k_rank_orig(int n, const int *orig, const int *scan, int *pos) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) pos[i] = scan[orig[i]]; }
static const int *out_positions(b200_sph *h)
{
  const int n = h->nlocal, m = std::max(h->next_orig, n) + 1;
  h->flag.ensure((size_t)m + 2); h->perm.ensure(n + 1); ensure_scan_tmp(h, m + 2);
  CK(cudaMemsetAsync(h->flag.p, 0, (size_t)(m + 1) * sizeof(int), h->st));
  LAUNCH(h, k_mark_orig, nblk(n, 256), 256, n, h->C().orig.p, h->flag.p);
  scan_exclusive(h, h->flag.p, m, h->scan_tmp.p);
  LAUNCH(h, k_rank_orig, nblk(n, 256), 256, n, h->C().orig.p, h->flag.p, h->perm.p);
  return h->perm.p;
}

static std::vector<int> out_positions_host(b200_sph *h)      // b200_get_neighbor_list (tests)
{
  std::vector<int> pos(h->nlocal);
  const int *src = h->world > 1 ? out_positions(h) : h->C().orig.p;
  CK(cudaStreamSynchronize(h->st));
  if (h->nlocal) CK(cudaMemcpy(pos.data(), src, (size_t)h->nlocal * sizeof(int), cudaMemcpyDeviceToHost));
  return pos;
}

int b200_get_atoms(b200_sph *h, int nmax, b200_atoms *a)
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  int n = h->nlocal;
  if (n > nmax) throw std::string("b200_get_atoms: buffer too small");
  if (!n) return 0;
  size_t nd, ni;
  HostMirror m = stage_layout(h, n, a, &nd, &ni);
  const int *outpos = nullptr;
  if (h->world > 1) outpos = out_positions(h);
  LAUNCH(h, k_unpack_atoms, nblk(n, 256), 256, n, m, pack_arrays(h), h->multiphase, outpos);
#define DOWN(dev, host, w) if (a->host) CK(cudaMemcpyAsync(a->host, m.dev, (size_t)(w) * n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  FOR_FIELDS(DOWN)
#undef DOWN
  if (a->type) CK(cudaMemcpyAsync(a->type, m.type, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  if (a->mask) CK(cudaMemcpyAsync(a->mask, m.mask, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  if (a->tag) CK(cudaMemcpyAsync(a->tag, m.tag, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  API_END
}

int b200_setup(b200_sph *h) { API_BEGIN CK(cudaSetDevice(h->device)); do_setup(h); API_END }
int b200_run(b200_sph *h, int nsteps) { API_BEGIN CK(cudaSetDevice(h->device)); do_run(h, nsteps); API_END }
int b200_initial_integrate(b200_sph *h) { API_BEGIN initial_integrate(h); API_END }
int b200_final_integrate(b200_sph *h) { API_BEGIN post_final(h, 0, 0, 1); API_END }
int b200_neigh_decide(b200_sph *h, int *rebuild) { API_BEGIN *rebuild = neigh_decide(h); API_END }
int b200_forward_comm(b200_sph *h) { API_BEGIN forward_comm(h); far_flags(h); API_END }
int b200_reneighbor(b200_sph *h) { API_BEGIN if (!h->geom_ready) setup_geometry(h); if (h->plan.empty()) build_plan(h); reneighbor(h); API_END }
int b200_force_clear(b200_sph *h) { API_BEGIN force_clear(h); API_END }
int b200_pair_compute(b200_sph *h, int slot)
{
  API_BEGIN
  if (slot < 0 || slot >= h->npair) throw std::string("b200_pair_compute: bad slot");
  Pass p{}; p.nslots = 1; p.slots[0] = slot;
  int st = h->h_tab[slot].style;
  p.type = st == B200_PAIR_RHOSUM ? 0 : st == B200_PAIR_RHOSUM_MULTIPHASE ? 1 : st == B200_PAIR_COLORGRADIENT ? 2 : 3;
  p.kinds = kind_of(st);
  run_pass(h, p);
  API_END
}
int b200_pair_compute_all(b200_sph *h) { API_BEGIN if (h->plan.empty()) build_plan(h); pair_compute_all(h); API_END }
int b200_reverse_comm(b200_sph *h) { API_BEGIN post_final(h, 1, 0, 0); API_END }
int b200_post_force(b200_sph *h) { API_BEGIN post_final(h, 0, 1, 0); API_END }

int b200_get_neighbor_list(b200_sph *h, int nlocal, int *numneigh, long long nentries, int *jtag, int *jimage)
{
  API_BEGIN
  CK(cudaSetDevice(h->device));
  if (nlocal != h->nlocal) throw std::string("b200_get_neighbor_list: nlocal mismatch");
  int n = nlocal, na = h->nall();
  if (!n) return 0;
  CK(cudaStreamSynchronize(h->st));
  std::vector<int> cnt(n), tag(na), img(na);
  std::vector<int> orig = out_positions_host(h);
  CK(cudaMemcpy(cnt.data(), h->numneigh.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> cfar(n);
  CK(cudaMemcpy(cfar.data(), h->numfar.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  if (h->rows_tiled) {        // slot lists of the tile path -> particle indices on the device, then as below
    for (int s = 0; s < n; s++) { cfar[s] = (cfar[s] & 0xffff) + (cfar[s] >> 16); numneigh[orig[s]] = cnt[s] + cfar[s]; }   // far + mid zone
    if (!jtag) return 0;
    long long tot = 0;
    std::vector<long long> off(n + 1);
    for (int i = 0; i < n; i++) { off[i] = tot; tot += numneigh[i]; }
    if (nentries < tot) throw std::string("b200_get_neighbor_list: buffer too small");
    CK(cudaMemcpy(tag.data(), h->C().tag.p, (size_t)na * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(img.data(), h->gimage.p, (size_t)na * sizeof(int), cudaMemcpyDeviceToHost));
    int width = 2 * h->stride;
    DevBuf<int> out; out.ensure((size_t)n * width);
    TileExportArgs X{n, h->stride / 8, width, h->multiphase ? (int)TMP_SLOT_MASK : (int)TILE_SLOT_MASK, h->gorder.p, h->tiles.p, h->ntiles, (const uint4 *)h->nbr.p, (const uint4 *)h->far.p, h->numneigh.p, h->numfar.p, out.p};
    LAUNCH(h, k_tile_export, std::max(1, std::min(h->ntiles, 1024)), 128, X);
    std::vector<int> rows((size_t)n * width);
    CK(cudaStreamSynchronize(h->st));
    CK(cudaMemcpy(rows.data(), out.p, rows.size() * sizeof(int), cudaMemcpyDeviceToHost));
    out.release();
    std::vector<std::pair<int, int>> tmp;
    for (int s = 0; s < n; s++) {
      tmp.clear();
      for (int k = 0; k < cnt[s] + cfar[s]; k++) { int j = rows[(size_t)s * width + k]; tmp.push_back({tag[j], img[j] & 0xff}); }
      std::sort(tmp.begin(), tmp.end());
      long long o = off[orig[s]];
      for (size_t k = 0; k < tmp.size(); k++) { jtag[o + k] = tmp[k].first; jimage[o + k] = tmp[k].second; }
    }
    return 0;
  }
  for (int s = 0; s < n; s++) numneigh[orig[s]] = (cnt[s] & 0xffff) + (cnt[s] >> 16) + cfar[s];
  if (!jtag) return 0;
  long long tot = 0;
  std::vector<long long> off(n + 1);
  for (int i = 0; i < n; i++) { off[i] = tot; tot += numneigh[i]; }
  if (nentries < tot) throw std::string("b200_get_neighbor_list: buffer too small");
  CK(cudaMemcpy(tag.data(), h->C().tag.p, (size_t)na * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(img.data(), h->gimage.p, (size_t)na * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<unsigned> rows((size_t)((n + 31) / 32) * 32 * h->stride);
  CK(cudaMemcpy(rows.data(), h->nbr.p, rows.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
  std::vector<unsigned> frows(rows.size());
  CK(cudaMemcpy(frows.data(), h->far.p, frows.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
  std::vector<std::pair<int, int>> tmp;
  for (int s = 0; s < n; s++) {
    tmp.clear();
    int nin = cnt[s] & 0xffff, nout = cnt[s] >> 16;
    for (int k = 0; k < nin + nout; k++) {
      int kk = k < nin ? k : h->stride - 1 - (k - nin);
      int j = rows[(size_t)(s >> 5) * h->stride * 32 + (size_t)kk * 32 + (s & 31)] & NBR_INDEX_MASK; tmp.push_back({tag[j], img[j] & 0xff});
    }
    for (int k = 0; k < cfar[s]; k++) {
      int j = frows[(size_t)(s >> 5) * h->stride * 32 + (size_t)k * 32 + (s & 31)] & NBR_INDEX_MASK; tmp.push_back({tag[j], img[j] & 0xff});
    }
    std::sort(tmp.begin(), tmp.end());
    long long o = off[orig[s]];
    for (size_t k = 0; k < tmp.size(); k++) { jtag[o + k] = tmp[k].first; jimage[o + k] = tmp[k].second; }
  }
  API_END
}

int b200_get_counters(b200_sph *h, long long c[8])
{ c[0] = h->launches; c[1] = h->nbuilds; c[2] = h->nsteps; c[3] = h->maxneigh; c[4] = h->nghost; c[5] = h->stride; c[6] = h->ninserted; c[7] = h->ndanger; return 0; }
int b200_set_timing(b200_sph *h, int on)
{ API_BEGIN h->tflush(); h->timing = on != 0; for (int k = 0; k < T_NTIMERS; k++) { h->t_ms[k] = 0; h->t_calls[k] = 0; } API_END }
int b200_get_timers(b200_sph *h, int n, double *ms, long long *calls)
{ API_BEGIN h->tflush(); for (int k = 0; k < n; k++) { ms[k] = k < T_NTIMERS ? h->t_ms[k] : 0.0; calls[k] = k < T_NTIMERS ? h->t_calls[k] : 0; } API_END }
const char *b200_timer_name(int i) { return (i >= 0 && i < T_NTIMERS) ? timer_names[i] : ""; }
int b200_sync(b200_sph *h) { API_BEGIN CK(cudaSetDevice(h->device)); CK(cudaStreamSynchronize(h->st)); API_END }

} // extern "C"
