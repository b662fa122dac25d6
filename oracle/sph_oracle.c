/* sph_oracle.c -- CPU restatement of the reference's SPH hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the oracle the CUDA engine is checked
 * against; it is never linked into, imported by, or called from the product
 * (libb200sph.so / the lammps-sph-multiphase_b200 package).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * Parity status: PINNED (one rank).  tests/test_oracle_golden.py checks this port
 * against (a) the six multiphase_two_atoms known-answer decks and (b) stage and
 * trajectory dumps of the real reference (oracle/_ref/liblammps_ref.so, built
 * from /root/reference by oracle/Makefile), committed under tests/golden/.
 * The P-rank emulation at the end of the file ("P ranks in one process") cannot
 * be run against the real reference here (the image has no MPI): it is pinned
 * through the decks whose result does not depend on the decomposition, which must
 * reproduce the 1-rank reference fixtures on 2-4 ranks (tests/test_world_cpu.py);
 * for decomposition-dependent decks at P > 1 (moving multiphase decks, fix phase_change
 * with one RanPark stream per rank, Atom::sort per rank) it is a restatement only,
 * held to its invariants (tags, mass) on the CPU.
 *
 * It is a deliberately plain, sequential, single-rank restatement that follows
 * the reference's own data structures (AoS per-atom arrays with ghosts behind
 * the owned atoms, linked-list bins, full list + derived half/skip filters,
 * accumulate-on-i-and-j with reverse communication), i.e. the opposite of the
 * CUDA engine's layout -- so agreement between the two is meaningful.
 * Every routine cites the reference file:line it restates (paths relative to
 * /root/reference/).  Compile with -ffp-contract=off (no FMA), like the
 * reference's generic x86-64 build.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sph_oracle.h"

#define MAXPAIR 16
#define MAXFIX 16
#define BIG 1.0e20
#define SMALL 1.0e-6 /* neighbor.cpp:44 */
#define EPSILON 1.0e-12 /* pair_sph_surfacetension.cpp */
#define CG_SMALL 1.0e-20 /* fix_phase_change.cpp:42 */

enum { FIX_MESO = 1, FIX_MESO_STATIONARY, FIX_GRAVITY, FIX_PHASE_CHANGE, FIX_SETMESO, FIX_ENFORCE2D, FIX_SETFORCE, FIX_SETMESODE, FIX_DT_RESET, FIX_ADDFORCE };

typedef struct {
  int style, nstep;
  int *mapped, *fixflag, *iskip;
  double *cut, *cutsq, *rho0, *B, *soundspeed, *gamma, *rbackground, *viscosity, *alpha, *tc;
} opair;

typedef struct {
  int kind, groupbit;
  double acc[3];
  int which, region_kind, match_inside; double value, region[6];
  int fset[3]; double fvalue[3];   /* fix setforce */
  char *formula[3];                /* fix addforce x y z / fix setmeso value: text of the v_ variable, or NULL */
  int pc_nins, pc_nlocal0;         /* fix phase_change between its local part and its finish: insertions of this call, nlocal before them */
  int nevery, minbound, maxbound; double tmin, tmax, xmax;   /* fix dt/reset */
  osph_phase_change_desc pc;
  long long next_reneighbor;
  int seed; /* RanPark state, random_park.cpp:22-47 */
} ofix;

struct osph_sph {
  int dim, periodic[3], multiphase, ntypes, ghost_velocity;
  double boxlo[3], boxhi[3], prd[3], sublo[3], subhi[3];
  int boundary[3][2], shrink; double small[3], minbox[3][2];   /* Domain::boundary / small / minxlo.. (domain.h:32,62,148) */
  double *mass;
  int nlocal, nghost, nmax;
  double *x, *v, *vest, *f, *cg;            /* [nmax][3] */
  double *rho, *drho, *e, *de, *cv, *rmass; /* [nmax] */
  int *type, *mask, *tag, *img;             /* img: ghost image code, 13 for owned */
  /* neighbor */
  double skin, *cutneighsq, cutneighmax, cutneighmaxsq, cutghost, triggersq;
  int every, delay, check, ago;
  long long ndanger, nbuilds;
  double *xhold; int maxhold;
  int nbinx, nbiny, nbinz, mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo, mbins;
  double bininvx, bininvy, bininvz, binsizex, binsizey, binsizez;
  int *binhead, *bins; int maxbin;
  int nstencil, *stencil;
  long long *firstneigh; int *numneigh; int *neigh; long long maxneigh; int maxlist;
  unsigned char *halfkeep; /* per full-list entry: kept by the reference's half list AT BUILD TIME (half_keeps) */
  int half_bin;            /* no full-list sub-style in the deck: the half list is half_bin_newton's, not half_from_full_newton's */
  /* comm: up to 6 swaps for P=1 (comm_brick.cpp:330-386) */
  int nswap, swapdim[6], swappbc[6], sendnum[6], firstrecv[6], *sendlist[6], maxsend[6];
  double slablo[6], slabhi[6];
  /* P > 1 emulated in one process (osph_world_*): this rank's place in the brick grid (comm.h procgrid/myloc/procneigh) */
  int inworld, me, procgrid[3], myloc[3], procneigh[3][2], sendproc[6], recvproc[6], dosend[6];
  /* styles */
  int npair; opair pair[MAXPAIR];
  int nfix; ofix fix[MAXFIX];
  int nold; ofix oldpc[MAXFIX];   /* fix phase_change entries of the registration osph_fix_clear dropped: an identical one keeps its state */
  double dt, ftm2v; long long ntimestep;
  double atime; long long atimestep, laststep;   /* Update::atime / atimestep, FixDtReset::laststep (osph_set_time / osph_get_time) */
  long long nsteps_done, ninserted, maxneighseen;
  int setup_done;
  int vir_request; double virial[6];
  /* atom_modify sort (atom.cpp:63-65, 540-552) */
  int sortfreq; double userbinsize; long long nextsort;
};

static char errbuf[512] = "";
static int fail(const char *msg) { snprintf(errbuf, sizeof errbuf, "%s", msg); return -1; }
const char *osph_last_error(void) { return errbuf; }
const char *osph_version(void) { return "sph_oracle 1 (CPU restatement; test infrastructure)"; }

/* ---------------------------------------------------------------------- */

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "oracle: out of memory\n"); abort(); } return q; }

static void grow(osph_sph *s, int n)
{
  if (n <= s->nmax) return;
  int old = s->nmax;
  int nm = n + n / 4 + 1024;
  s->x = xrealloc(s->x, sizeof(double) * 3 * nm);   s->v = xrealloc(s->v, sizeof(double) * 3 * nm);
  s->vest = xrealloc(s->vest, sizeof(double) * 3 * nm); s->f = xrealloc(s->f, sizeof(double) * 3 * nm);
  s->cg = xrealloc(s->cg, sizeof(double) * 3 * nm);
  s->rho = xrealloc(s->rho, sizeof(double) * nm);   s->drho = xrealloc(s->drho, sizeof(double) * nm);
  s->e = xrealloc(s->e, sizeof(double) * nm);       s->de = xrealloc(s->de, sizeof(double) * nm);
  s->cv = xrealloc(s->cv, sizeof(double) * nm);     s->rmass = xrealloc(s->rmass, sizeof(double) * nm);
  s->type = xrealloc(s->type, sizeof(int) * nm);    s->mask = xrealloc(s->mask, sizeof(int) * nm);
  s->tag = xrealloc(s->tag, sizeof(int) * nm);      s->img = xrealloc(s->img, sizeof(int) * nm);
  /* fresh memory reads as zero (the reference's arrays come from malloc of fresh pages) */
  for (int i = old; i < nm; i++) {
    for (int d = 0; d < 3; d++) s->x[3*i+d] = s->v[3*i+d] = s->vest[3*i+d] = s->f[3*i+d] = s->cg[3*i+d] = 0.0;
    s->rho[i] = s->drho[i] = s->e[i] = s->de[i] = s->cv[i] = s->rmass[i] = 0.0;
    s->type[i] = s->mask[i] = s->tag[i] = 0; s->img[i] = 13;
  }
  s->nmax = nm;
}

int osph_create(osph_sph **h, int device)
{
  (void)device;
  osph_sph *s = calloc(1, sizeof *s);
  if (!s) return fail("calloc");
  s->dim = 3; s->ftm2v = 1.0; s->every = 1; s->delay = 10; s->check = 1;
  *h = s;
  return 0;
}

static void free_pair(opair *p)
{
  free(p->mapped); free(p->fixflag); free(p->iskip); free(p->cut); free(p->cutsq); free(p->rho0); free(p->B);
  free(p->soundspeed); free(p->gamma); free(p->rbackground); free(p->viscosity); free(p->alpha); free(p->tc);
  memset(p, 0, sizeof *p);
}

int osph_destroy(osph_sph *s)
{
  if (!s) return 0;
  for (int i = 0; i < s->npair; i++) free_pair(&s->pair[i]);
  for (int i = 0; i < 6; i++) free(s->sendlist[i]);
  free(s->x); free(s->v); free(s->vest); free(s->f); free(s->cg); free(s->rho); free(s->drho); free(s->e); free(s->de);
  free(s->cv); free(s->rmass); free(s->type); free(s->mask); free(s->tag); free(s->img); free(s->mass);
  free(s->cutneighsq); free(s->xhold); free(s->binhead); free(s->bins); free(s->stencil);
  free(s->firstneigh); free(s->numneigh); free(s->neigh); free(s->halfkeep);
  free(s);
  return 0;
}

int osph_comm_unique_id(char id[128]) { memset(id, 0, 128); return 0; }
/* world > 1: the instance becomes one rank of an emulated brick decomposition; it can then only be driven through
 * osph_world_setup / osph_world_run together with its peers (oracle-only entry points, sph_oracle.h) */
int osph_comm_init(osph_sph *s, int world, int rank, const int procgrid[3], const int myloc[3], const int procneigh[6], const char id[128])
{
  (void)id;
  if (world == 1) return 0;
  if (procgrid[0] * procgrid[1] * procgrid[2] != world) return fail("comm_init: procgrid does not match the world size");
  s->inworld = 1; s->me = rank;
  for (int d = 0; d < 3; d++) { s->procgrid[d] = procgrid[d]; s->myloc[d] = myloc[d]; s->procneigh[d][0] = procneigh[2*d]; s->procneigh[d][1] = procneigh[2*d+1]; }
  return 0;
}

int osph_domain(osph_sph *s, int dim, const double boxlo[3], const double boxhi[3], const int periodicity[3],
                const double sublo[3], const double subhi[3])
{
  if (dim != 2 && dim != 3) return fail("dimension must be 2 or 3");
  s->dim = dim;
  for (int d = 0; d < 3; d++) {
    s->boxlo[d] = boxlo[d]; s->boxhi[d] = boxhi[d]; s->prd[d] = boxhi[d] - boxlo[d]; /* domain.cpp set_global_box */
    s->periodic[d] = periodicity[d];
    s->sublo[d] = sublo ? sublo[d] : boxlo[d]; s->subhi[d] = subhi ? subhi[d] : boxhi[d];
    if ((s->sublo[d] != boxlo[d] || s->subhi[d] != boxhi[d]) && !s->inworld) return fail("oracle: a sub-domain smaller than the box needs osph_comm_init (world mode) first");
  }
  return 0;
}

/* boundary s / m: Domain::set_boundary (domain.cpp:1486-1491, nonperiodic = 2) */
int osph_boundary(osph_sph *s, const int boundary[6], const double small[3], const double minbox[6])
{
  s->shrink = 0;
  for (int d = 0; d < 3; d++) {
    for (int k = 0; k < 2; k++) {
      int b = boundary[2*d+k];
      if (b < 0 || b > 3) return fail("boundary style must be 0 (p), 1 (f), 2 (s) or 3 (m)");
      if ((b == 0) != (s->periodic[d] != 0)) return fail("boundary styles disagree with the periodicity given to b200_domain");
      s->boundary[d][k] = b; s->minbox[d][k] = minbox ? minbox[2*d+k] : 0.0;
      if (b >= 2) s->shrink = 1;
    }
    s->small[d] = small ? small[d] : 0.0;
  }
  return 0;
}
int osph_get_box(osph_sph *s, double boxlo[3], double boxhi[3])
{ for (int d = 0; d < 3; d++) { boxlo[d] = s->boxlo[d]; boxhi[d] = s->boxhi[d]; } return 0; }

int osph_atom_style(osph_sph *s, int multiphase, int ntypes, const double *mass)
{
  s->multiphase = multiphase; s->ntypes = ntypes;
  s->mass = xrealloc(s->mass, sizeof(double) * (ntypes + 1));
  for (int i = 0; i <= ntypes; i++) s->mass[i] = mass ? mass[i] : 0.0;
  return 0;
}

int osph_neighbor(osph_sph *s, double skin, int every, int delay, int check, const double *cutneighsq,
                  double cutneighmax, double cutghost)
{
  int n = s->ntypes + 1;
  s->skin = skin; s->every = every; s->delay = delay; s->check = check;
  s->cutneighsq = xrealloc(s->cutneighsq, sizeof(double) * n * n);
  memcpy(s->cutneighsq, cutneighsq, sizeof(double) * n * n);
  s->cutneighmax = cutneighmax; s->cutneighmaxsq = cutneighmax * cutneighmax; /* neighbor.cpp:282 */
  s->cutghost = cutghost;
  s->triggersq = 0.25 * skin * skin; /* neighbor.cpp:240 */
  return 0;
}

int osph_timestep(osph_sph *s, double dt, double ftm2v, long long ntimestep)
{ s->dt = dt; s->ftm2v = ftm2v; s->ntimestep = ntimestep; return 0; }

int osph_comm_modify(osph_sph *s, int ghost_velocity) { s->ghost_velocity = ghost_velocity; return 0; }
int osph_atom_modify(osph_sph *s, int sortfreq, double userbinsize)
{
  if (sortfreq < 0 || userbinsize < 0.0) return fail("Illegal atom_modify command");
  s->sortfreq = sortfreq; s->userbinsize = userbinsize;
  return 0;
}

int osph_pair_clear(osph_sph *s) { for (int i = 0; i < s->npair; i++) free_pair(&s->pair[i]); s->npair = 0; return 0; }

static char *dups(const char *p) { char *q = malloc(strlen(p) + 1); strcpy(q, p); return q; }
static double *dupd(const double *p, int n) { if (!p) return NULL; double *q = malloc(sizeof(double) * n); memcpy(q, p, sizeof(double) * n); return q; }
static int *dupi(const int *p, int n) { if (!p) return NULL; int *q = malloc(sizeof(int) * n); memcpy(q, p, sizeof(int) * n); return q; }

int osph_pair_add(osph_sph *s, const osph_pair_desc *d)
{
  if (s->npair == MAXPAIR) return fail("too many pair sub-styles");
  int n = s->ntypes + 1, nn = n * n;
  if (!d->mapped || !d->cut || !d->cutsq) return fail("pair desc needs mapped, cut, cutsq");
  opair *p = &s->pair[s->npair];
  memset(p, 0, sizeof *p);
  p->style = d->style; p->nstep = d->nstep;
  p->mapped = dupi(d->mapped, nn); p->cut = dupd(d->cut, nn); p->cutsq = dupd(d->cutsq, nn);
  p->rho0 = dupd(d->rho0, n); p->B = dupd(d->B, n); p->soundspeed = dupd(d->soundspeed, n);
  p->gamma = dupd(d->gamma, n); p->rbackground = dupd(d->rbackground, n);
  p->viscosity = dupd(d->viscosity, nn); p->alpha = dupd(d->alpha, nn); p->tc = dupd(d->tc, nn);
  p->fixflag = dupi(d->fixflag, nn);
  /* iskip: skip I entirely if no J is mapped (pair_hybrid.cpp:473-477) */
  p->iskip = calloc(n, sizeof(int));
  for (int i = 1; i < n; i++) { p->iskip[i] = 1; for (int j = 1; j < n; j++) if (p->mapped[i * n + j]) p->iskip[i] = 0; }
  return s->npair++;
}

/* FixPhaseChange keeps next_reneighbor and its RanPark position across `run` commands (fix_phase_change.cpp:116,345); the caller
 * re-registers the deck at every setup, so an identical fix phase_change takes its state over from the previous registration */
int osph_fix_clear(osph_sph *s)
{
  s->nold = 0;
  for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == FIX_PHASE_CHANGE) s->oldpc[s->nold++] = s->fix[i];
  s->nfix = 0;
  return 0;
}
static ofix *newfix(osph_sph *s, int kind, int groupbit)
{ if (s->nfix == MAXFIX) return NULL; ofix *f = &s->fix[s->nfix++]; memset(f, 0, sizeof *f); f->kind = kind; f->groupbit = groupbit; return f; }
int osph_fix_meso(osph_sph *s, int groupbit) { return newfix(s, FIX_MESO, groupbit) ? 0 : fail("too many fixes"); }
int osph_fix_meso_stationary(osph_sph *s, int groupbit) { return newfix(s, FIX_MESO_STATIONARY, groupbit) ? 0 : fail("too many fixes"); }
int osph_fix_gravity(osph_sph *s, int groupbit, double xacc, double yacc, double zacc)
{ ofix *f = newfix(s, FIX_GRAVITY, groupbit); if (!f) return fail("too many fixes"); f->acc[0] = xacc; f->acc[1] = yacc; f->acc[2] = zacc; return 0; }
int osph_fix_setmeso(osph_sph *s, int groupbit, int which, double value, int region_kind, const double region[6], int match_inside)
{
  ofix *f = newfix(s, FIX_SETMESO, groupbit); if (!f) return fail("too many fixes");
  f->which = which; f->value = value; f->region_kind = region_kind; f->match_inside = match_inside;
  for (int q = 0; q < 6; q++) f->region[q] = (region_kind && region) ? region[q] : 0.0;
  return 0;
}
int osph_fix_setmeso_var(osph_sph *s, int groupbit, int which, const char *formula, int region_kind, const double region[6], int match_inside)
{
  if (!formula) return fail("osph_fix_setmeso_var: no formula");
  if (osph_fix_setmeso(s, groupbit, which, 0.0, region_kind, region, match_inside)) return -1;
  s->fix[s->nfix - 1].formula[0] = dups(formula);
  return 0;
}
int osph_fix_addforce(osph_sph *s, int groupbit, const double value[3], const char *const formula[3])
{
  ofix *f = newfix(s, FIX_ADDFORCE, groupbit); if (!f) return fail("too many fixes");
  for (int d = 0; d < 3; d++) { f->acc[d] = value ? value[d] : 0.0; f->formula[d] = (formula && formula[d]) ? dups(formula[d]) : NULL; }
  return 0;
}
int osph_fix_enforce2d(osph_sph *s, int groupbit) { return newfix(s, FIX_ENFORCE2D, groupbit) ? 0 : fail("too many fixes"); }
int osph_fix_dt_reset(osph_sph *s, int groupbit, int nevery, int minbound, double tmin, int maxbound, double tmax, double xmax)
{
  if (nevery <= 0 || xmax <= 0.0 || (minbound && tmin < 0.0) || (maxbound && tmax < 0.0) || (minbound && maxbound && tmin >= tmax))
    return fail("Illegal fix dt/reset command");
  ofix *f = newfix(s, FIX_DT_RESET, groupbit);
  if (!f) return fail("too many fixes");
  f->nevery = nevery; f->minbound = minbound; f->tmin = tmin; f->maxbound = maxbound; f->tmax = tmax; f->xmax = xmax;
  return 0;
}
int osph_get_timestep(osph_sph *s, double *dt) { *dt = s->dt; return 0; }
int osph_set_time(osph_sph *s, double atime, long long atimestep, long long laststep) { s->atime = atime; s->atimestep = atimestep; s->laststep = laststep; return 0; }
int osph_get_time(osph_sph *s, double *atime, long long *atimestep, long long *laststep)
{ if (atime) *atime = s->atime; if (atimestep) *atimestep = s->atimestep; if (laststep) *laststep = s->laststep; return 0; }
int osph_fix_setmesode(osph_sph *s, int groupbit, double value, int region_kind, const double region[6])
{
  ofix *f = newfix(s, FIX_SETMESODE, groupbit);
  if (!f) return fail("too many fixes");
  f->value = value; f->region_kind = region_kind; f->match_inside = 1;
  for (int q = 0; q < 6; q++) f->region[q] = (region_kind && region) ? region[q] : 0.0;
  return 0;
}
int osph_fix_setforce(osph_sph *s, int groupbit, const int set[3], const double value[3])
{
  ofix *f = newfix(s, FIX_SETFORCE, groupbit);
  if (!f) return fail("too many fixes");
  for (int d = 0; d < 3; d++) { f->fset[d] = set[d] != 0; f->fvalue[d] = value[d]; }
  return 0;
}
int osph_fix_phase_change(osph_sph *s, const osph_phase_change_desc *d)
{
  if (d->seed <= 0) return fail("Illegal value for seed"); /* fix_phase_change.cpp:70 */
  ofix *f = newfix(s, FIX_PHASE_CHANGE, d->groupbit); if (!f) return fail("too many fixes");
  f->pc = *d; f->next_reneighbor = d->first_step; f->seed = d->seed;
  for (int k = 0; k < s->nold; k++) {
    const osph_phase_change_desc *o = &s->oldpc[k].pc;
    if (s->oldpc[k].groupbit == d->groupbit && o->groupbit == d->groupbit && o->Tc == d->Tc && o->Tt == d->Tt && o->Hwv == d->Hwv && o->dr == d->dr && o->to_mass == d->to_mass &&
        o->cutoff == d->cutoff && o->from_type == d->from_type && o->to_type == d->to_type && o->nfreq == d->nfreq && o->seed == d->seed &&
        o->energy_chance_flag == d->energy_chance_flag && o->change_chance == d->change_chance && o->phase_change_rate == d->phase_change_rate &&
        o->maxattempt == d->maxattempt && o->first_step == d->first_step) {
      f->next_reneighbor = s->oldpc[k].next_reneighbor; f->seed = s->oldpc[k].seed;
      s->oldpc[k] = s->oldpc[--s->nold];
      break;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------- */

int osph_set_atoms(osph_sph *s, int n, const osph_atoms *a)
{
  if (!a->x || !a->type) return fail("set_atoms needs x and type");
  grow(s, n);
  s->nlocal = n; s->nghost = 0;
  for (int i = 0; i < n; i++) {
    for (int d = 0; d < 3; d++) {
      s->x[3*i+d] = a->x[3*i+d];
      s->v[3*i+d] = a->v ? a->v[3*i+d] : 0.0;
      s->vest[3*i+d] = a->vest ? a->vest[3*i+d] : s->v[3*i+d];
      s->f[3*i+d] = a->f ? a->f[3*i+d] : 0.0;
      s->cg[3*i+d] = a->colorgradient ? a->colorgradient[3*i+d] : 0.0;
    }
    s->type[i] = a->type[i]; s->mask[i] = a->mask ? a->mask[i] : 1; s->tag[i] = a->tag ? a->tag[i] : i + 1; s->img[i] = 13;
    s->rho[i] = a->rho ? a->rho[i] : 0.0; s->drho[i] = a->drho ? a->drho[i] : 0.0;
    s->e[i] = a->e ? a->e[i] : 0.0; s->de[i] = a->de ? a->de[i] : 0.0; s->cv[i] = a->cv ? a->cv[i] : 0.0;
    s->rmass[i] = a->rmass ? a->rmass[i] : (s->mass ? s->mass[s->type[i]] : 0.0);
  }
  s->setup_done = 0;
  return 0;
}

int osph_get_natoms(osph_sph *s, int *nlocal, int *nghost) { if (nlocal) *nlocal = s->nlocal; if (nghost) *nghost = s->nghost; return 0; }

static int copy_out(osph_sph *s, int n, int nmax, osph_atoms *a)
{
  if (n > nmax) return fail("get_atoms: buffer too small");
  if (a->x) memcpy(a->x, s->x, sizeof(double) * 3 * n);
  if (a->v) memcpy(a->v, s->v, sizeof(double) * 3 * n);
  if (a->vest) memcpy(a->vest, s->vest, sizeof(double) * 3 * n);
  if (a->f) memcpy(a->f, s->f, sizeof(double) * 3 * n);
  if (a->colorgradient) memcpy(a->colorgradient, s->cg, sizeof(double) * 3 * n);
  if (a->rho) memcpy(a->rho, s->rho, sizeof(double) * n);
  if (a->drho) memcpy(a->drho, s->drho, sizeof(double) * n);
  if (a->e) memcpy(a->e, s->e, sizeof(double) * n);
  if (a->de) memcpy(a->de, s->de, sizeof(double) * n);
  if (a->cv) memcpy(a->cv, s->cv, sizeof(double) * n);
  if (a->rmass) memcpy(a->rmass, s->rmass, sizeof(double) * n);
  if (a->type) memcpy(a->type, s->type, sizeof(int) * n);
  if (a->mask) memcpy(a->mask, s->mask, sizeof(int) * n);
  if (a->tag) memcpy(a->tag, s->tag, sizeof(int) * n);
  return 0;
}
int osph_get_atoms(osph_sph *s, int nmax, osph_atoms *a) { return copy_out(s, s->nlocal, nmax, a); }
int osph_get_all(osph_sph *s, int nmax, osph_atoms *a) { return copy_out(s, s->nlocal + s->nghost, nmax, a); }

/* ======================================================================
   domain / comm
   ====================================================================== */

/* Domain::pbc, src/domain.cpp:476-560 (orthogonal box, no deform) */
static void domain_pbc(osph_sph *s)
{
  for (int i = 0; i < s->nlocal; i++)
    for (int d = 0; d < 3; d++) {
      if (!s->periodic[d]) continue;
      double *xc = &s->x[3*i+d];
      if (*xc < s->boxlo[d]) *xc += s->prd[d];
      if (*xc >= s->boxhi[d]) { *xc -= s->prd[d]; if (*xc < s->boxlo[d]) *xc = s->boxlo[d]; /* MAX(x,lo) */ }
    }
}

/* Domain::reset_box, src/domain.cpp:338-406 (orthogonal box): shrink-wrapped faces follow the extent of the owned atoms,
 * then set_global_box / set_local_box (:268-330) refresh prd and the (single-rank) sub-domain. */
static int domain_reset_box(osph_sph *s)
{
  if (!s->shrink) return 0;
  double lo[3] = {BIG, BIG, BIG}, hi[3] = {-BIG, -BIG, -BIG};
  for (int i = 0; i < s->nlocal; i++)
    for (int d = 0; d < 3; d++) { double c = s->x[3*i+d]; if (c < lo[d]) lo[d] = c; if (c > hi[d]) hi[d] = c; }
  for (int d = 0; d < 3; d++) {
    if (s->periodic[d]) continue;
    if (s->boundary[d][0] == 2) s->boxlo[d] = lo[d] - s->small[d];
    else if (s->boundary[d][0] == 3) { double c = lo[d] - s->small[d]; s->boxlo[d] = c < s->minbox[d][0] ? c : s->minbox[d][0]; }
    if (s->boundary[d][1] == 2) s->boxhi[d] = hi[d] + s->small[d];
    else if (s->boundary[d][1] == 3) { double c = hi[d] + s->small[d]; s->boxhi[d] = c > s->minbox[d][1] ? c : s->minbox[d][1]; }
    if (s->boxlo[d] > s->boxhi[d]) return fail("Illegal simulation box");
  }
  for (int d = 0; d < 3; d++) { s->prd[d] = s->boxhi[d] - s->boxlo[d]; s->sublo[d] = s->boxlo[d]; s->subhi[d] = s->boxhi[d]; }
  return 0;
}

/* CommBrick::setup for a 1x1x1 processor grid, comm_brick.cpp:150-386.
 * maxneed[d] = int(cutghost*1/prd)+1, 0 for non-periodic dims (MIN(maxneed,procgrid-1))
 * and for z in 2-D.  Only maxneed<=1 is restated (cutghost < prd).             */
static int comm_setup(osph_sph *s)
{
  s->nswap = 0;
  for (int d = 0; d < 3; d++) {
    int maxneed = (int)(s->cutghost * 1 / s->prd[d]) + 1; /* :228-230 */
    if (s->dim == 2 && d == 2) maxneed = 0;
    if (!s->periodic[d]) maxneed = 0; /* MIN(maxneed, procgrid-1) */
    if (maxneed > 1) return fail("oracle: cutghost >= box length in a periodic dimension is not restated");
    for (int ineed = 0; ineed < 2 * maxneed; ineed++) {
      int k = s->nswap++;
      s->swapdim[k] = d;
      if (ineed % 2 == 0) { s->slablo[k] = -BIG; s->slabhi[k] = s->sublo[d] + s->cutghost; s->swappbc[k] = 1; }  /* :340-351 */
      else                { s->slablo[k] = s->subhi[d] - s->cutghost; s->slabhi[k] = BIG; s->swappbc[k] = -1; } /* :361-374 */
    }
  }
  return 0;
}

/* CommBrick::borders, comm_brick.cpp:696-864, with AtomVecMeso*::pack_border[_vel]
 * (atom_vec_meso_multiphase.cpp:555-721, atom_vec_meso.cpp pack_border) on a
 * self-swap (sendproc == me: buf = buf_send).                                   */
static void comm_borders(osph_sph *s)
{
  s->nghost = 0;
  int iswap = 0;
  while (iswap < s->nswap) {
    int dim = s->swapdim[iswap];
    int nfirst = 0, nlast = s->nlocal + s->nghost; /* both swaps of a dim scan the same range (:722-725) */
    for (int half = 0; half < 2 && iswap < s->nswap && s->swapdim[iswap] == dim; half++, iswap++) {
      double lo = s->slablo[iswap], hi = s->slabhi[iswap];
      int nsend = 0;
      for (int i = nfirst; i < nlast; i++)
        if (s->x[3*i+dim] >= lo && s->x[3*i+dim] <= hi) {
          if (nsend == s->maxsend[iswap]) { s->maxsend[iswap] = nsend * 2 + 1024; s->sendlist[iswap] = xrealloc(s->sendlist[iswap], sizeof(int) * s->maxsend[iswap]); }
          s->sendlist[iswap][nsend++] = i;
        }
      int first = s->nlocal + s->nghost;
      grow(s, first + nsend);
      double shift = s->swappbc[iswap] * s->prd[dim]; /* dx = pbc[0]*domain->xprd */
      int imgmul = dim == 0 ? 1 : (dim == 1 ? 3 : 9);
      for (int k = 0; k < nsend; k++) {
        int j = s->sendlist[iswap][k], g = first + k;
        for (int d = 0; d < 3; d++) {
          s->x[3*g+d] = (d == dim) ? s->x[3*j+d] + shift : s->x[3*j+d];
          s->cg[3*g+d] = s->cg[3*j+d]; s->vest[3*g+d] = s->vest[3*j+d];
          if (s->ghost_velocity) s->v[3*g+d] = s->v[3*j+d];
        }
        s->tag[g] = s->tag[j]; s->type[g] = s->type[j]; s->mask[g] = s->mask[j];
        s->rho[g] = s->rho[j]; s->rmass[g] = s->rmass[j]; s->e[g] = s->e[j]; s->cv[g] = s->cv[j];
        s->img[g] = s->img[j] + s->swappbc[iswap] * imgmul;
      }
      s->sendnum[iswap] = nsend; s->firstrecv[iswap] = first; s->nghost += nsend;
    }
  }
}

/* CommBrick::forward_comm, comm_brick.cpp:444-506 with pack_comm[_vel]
 * (atom_vec_meso_multiphase.cpp:319-465: x,rho,cg,rmass,e,vest(+v);
 *  atom_vec_meso.cpp:246-270: x,rho,e,vest(+v)); cv/type/tag/mask are NOT resent */
static void comm_forward(osph_sph *s)
{
  for (int iswap = 0; iswap < s->nswap; iswap++) {
    int dim = s->swapdim[iswap];
    double shift = s->swappbc[iswap] * s->prd[dim];
    for (int k = 0; k < s->sendnum[iswap]; k++) {
      int j = s->sendlist[iswap][k], g = s->firstrecv[iswap] + k;
      for (int d = 0; d < 3; d++) {
        s->x[3*g+d] = (d == dim) ? s->x[3*j+d] + shift : s->x[3*j+d];
        s->vest[3*g+d] = s->vest[3*j+d];
        if (s->multiphase) s->cg[3*g+d] = s->cg[3*j+d];
        if (s->ghost_velocity) s->v[3*g+d] = s->v[3*j+d];
      }
      s->rho[g] = s->rho[j]; s->e[g] = s->e[j];
      if (s->multiphase) s->rmass[g] = s->rmass[j];
    }
  }
}

/* CommBrick::reverse_comm, comm_brick.cpp:513-560, pack_reverse = f3,drho,de
 * (atom_vec_meso_multiphase.cpp:520-551); swaps in reverse order                */
static void comm_reverse(osph_sph *s)
{
  for (int iswap = s->nswap - 1; iswap >= 0; iswap--)
    for (int k = 0; k < s->sendnum[iswap]; k++) {
      int j = s->sendlist[iswap][k], g = s->firstrecv[iswap] + k;
      for (int d = 0; d < 3; d++) s->f[3*j+d] += s->f[3*g+d];
      s->drho[j] += s->drho[g]; s->de[j] += s->de[g];
    }
}

/* CommBrick::forward_comm_pair with PairSPHRhoSum::pack_forward_comm (rho),
 * comm_brick.cpp:871-904, pair_sph_rhosum.cpp:290-313                            */
static void comm_forward_rho(osph_sph *s)
{
  for (int iswap = 0; iswap < s->nswap; iswap++)
    for (int k = 0; k < s->sendnum[iswap]; k++) s->rho[s->firstrecv[iswap] + k] = s->rho[s->sendlist[iswap][k]];
}

/* CommBrick::reverse_comm_fix with FixPhaseChange::pack/unpack_reverse_comm (dmass == drho),
 * comm_brick.cpp:999-1037, fix_phase_change.cpp:515-538                           */
static void comm_reverse_dmass(osph_sph *s)
{
  for (int iswap = s->nswap - 1; iswap >= 0; iswap--)
    for (int k = 0; k < s->sendnum[iswap]; k++) s->drho[s->sendlist[iswap][k]] += s->drho[s->firstrecv[iswap] + k];
}

/* ======================================================================
   neighbor
   ====================================================================== */

/* Neighbor::bin_distance, neighbor.cpp:1751-1775 */
static double bin_distance(osph_sph *s, int i, int j, int k)
{
  double delx, dely, delz;
  if (i > 0) delx = (i - 1) * s->binsizex; else if (i == 0) delx = 0.0; else delx = (i + 1) * s->binsizex;
  if (j > 0) dely = (j - 1) * s->binsizey; else if (j == 0) dely = 0.0; else dely = (j + 1) * s->binsizey;
  if (k > 0) delz = (k - 1) * s->binsizez; else if (k == 0) delz = 0.0; else delz = (k + 1) * s->binsizez;
  return (delx * delx + dely * dely + delz * delz);
}

/* Neighbor::setup_bins, neighbor.cpp:1581-1745 + stencil_full_bin_{2d,3d}, neigh_stencil.cpp:395-448 */
static int setup_bins(osph_sph *s)
{
  double bbox[3], bsubboxlo[3], bsubboxhi[3];
  for (int d = 0; d < 3; d++) { bsubboxlo[d] = s->sublo[d] - s->cutghost; bsubboxhi[d] = s->subhi[d] + s->cutghost; bbox[d] = s->boxhi[d] - s->boxlo[d]; }
  double binsize_optimal = 0.5 * s->cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  double binsizeinv = 1.0 / binsize_optimal;
  s->nbinx = (int)(bbox[0] * binsizeinv); s->nbiny = (int)(bbox[1] * binsizeinv);
  s->nbinz = s->dim == 3 ? (int)(bbox[2] * binsizeinv) : 1;
  if (s->nbinx == 0) s->nbinx = 1; if (s->nbiny == 0) s->nbiny = 1; if (s->nbinz == 0) s->nbinz = 1;
  s->binsizex = bbox[0] / s->nbinx; s->binsizey = bbox[1] / s->nbiny; s->binsizez = bbox[2] / s->nbinz;
  s->bininvx = 1.0 / s->binsizex; s->bininvy = 1.0 / s->binsizey; s->bininvz = 1.0 / s->binsizez;
  int mbinxhi, mbinyhi, mbinzhi = 0; double coord;
  coord = bsubboxlo[0] - SMALL * bbox[0]; s->mbinxlo = (int)((coord - s->boxlo[0]) * s->bininvx); if (coord < s->boxlo[0]) s->mbinxlo--;
  coord = bsubboxhi[0] + SMALL * bbox[0]; mbinxhi = (int)((coord - s->boxlo[0]) * s->bininvx);
  coord = bsubboxlo[1] - SMALL * bbox[1]; s->mbinylo = (int)((coord - s->boxlo[1]) * s->bininvy); if (coord < s->boxlo[1]) s->mbinylo--;
  coord = bsubboxhi[1] + SMALL * bbox[1]; mbinyhi = (int)((coord - s->boxlo[1]) * s->bininvy);
  if (s->dim == 3) {
    coord = bsubboxlo[2] - SMALL * bbox[2]; s->mbinzlo = (int)((coord - s->boxlo[2]) * s->bininvz); if (coord < s->boxlo[2]) s->mbinzlo--;
    coord = bsubboxhi[2] + SMALL * bbox[2]; mbinzhi = (int)((coord - s->boxlo[2]) * s->bininvz);
  }
  s->mbinxlo--; mbinxhi++; s->mbinx = mbinxhi - s->mbinxlo + 1;
  s->mbinylo--; mbinyhi++; s->mbiny = mbinyhi - s->mbinylo + 1;
  if (s->dim == 3) { s->mbinzlo--; mbinzhi++; } else s->mbinzlo = mbinzhi = 0;
  s->mbinz = mbinzhi - s->mbinzlo + 1;
  s->mbins = s->mbinx * s->mbiny * s->mbinz;
  s->binhead = xrealloc(s->binhead, sizeof(int) * s->mbins);
  int sx = (int)(s->cutneighmax * s->bininvx); if (sx * s->binsizex < s->cutneighmax) sx++;
  int sy = (int)(s->cutneighmax * s->bininvy); if (sy * s->binsizey < s->cutneighmax) sy++;
  int sz = (int)(s->cutneighmax * s->bininvz); if (sz * s->binsizez < s->cutneighmax) sz++;
  if (s->dim == 2) sz = 0;
  s->stencil = xrealloc(s->stencil, sizeof(int) * (2*sx+1) * (2*sy+1) * (2*sz+1));
  s->nstencil = 0;
  for (int k = -sz; k <= sz; k++) for (int j = -sy; j <= sy; j++) for (int i = -sx; i <= sx; i++)
    if (bin_distance(s, i, j, k) < s->cutneighmaxsq) s->stencil[s->nstencil++] = k * s->mbiny * s->mbinx + j * s->mbinx + i;
  return 0;
}

/* Neighbor::coord2bin, neighbor.cpp:1961-1990 */
static void coord2bin3(osph_sph *s, const double *x, int *pix, int *piy, int *piz)
{
  int ix, iy, iz;
  if (x[0] >= s->boxhi[0]) ix = (int)((x[0] - s->boxhi[0]) * s->bininvx) + s->nbinx;
  else if (x[0] >= s->boxlo[0]) { ix = (int)((x[0] - s->boxlo[0]) * s->bininvx); if (ix > s->nbinx - 1) ix = s->nbinx - 1; }
  else ix = (int)((x[0] - s->boxlo[0]) * s->bininvx) - 1;
  if (x[1] >= s->boxhi[1]) iy = (int)((x[1] - s->boxhi[1]) * s->bininvy) + s->nbiny;
  else if (x[1] >= s->boxlo[1]) { iy = (int)((x[1] - s->boxlo[1]) * s->bininvy); if (iy > s->nbiny - 1) iy = s->nbiny - 1; }
  else iy = (int)((x[1] - s->boxlo[1]) * s->bininvy) - 1;
  if (x[2] >= s->boxhi[2]) iz = (int)((x[2] - s->boxhi[2]) * s->bininvz) + s->nbinz;
  else if (x[2] >= s->boxlo[2]) { iz = (int)((x[2] - s->boxlo[2]) * s->bininvz); if (iz > s->nbinz - 1) iz = s->nbinz - 1; }
  else iz = (int)((x[2] - s->boxlo[2]) * s->bininvz) - 1;
  *pix = ix; *piy = iy; *piz = iz;
}
static int coord2bin(osph_sph *s, const double *x)
{
  int ix, iy, iz;
  coord2bin3(s, x, &ix, &iy, &iz);
  return (iz - s->mbinzlo) * s->mbiny * s->mbinx + (iy - s->mbinylo) * s->mbinx + (ix - s->mbinxlo);
}

static int half_keeps(osph_sph *s, int i, int j);

/* Neighbor::build (neighbor.cpp:1418-1499) -> bin_atoms (:1911-1947) + full_bin (neigh_full.cpp:241-340) */
static int neighbor_build(osph_sph *s)
{
  int nlocal = s->nlocal, nall = nlocal + s->nghost, nt = s->ntypes + 1;
  s->ago = 0; s->nbuilds++;
  s->half_bin = 1;
  for (int k = 0; k < s->npair; k++)
    if (s->pair[k].style == B200_PAIR_RHOSUM || s->pair[k].style == B200_PAIR_RHOSUM_MULTIPHASE || s->pair[k].style == B200_PAIR_COLORGRADIENT) s->half_bin = 0;
  if (s->check) { /* dist_check: store xhold (:1428-1441) */
    if (nlocal > s->maxhold) { s->maxhold = s->nmax; s->xhold = xrealloc(s->xhold, sizeof(double) * 3 * s->maxhold); }
    memcpy(s->xhold, s->x, sizeof(double) * 3 * nlocal);
  }
  if (s->nmax > s->maxbin) { s->maxbin = s->nmax; s->bins = xrealloc(s->bins, sizeof(int) * s->maxbin); }
  for (int i = 0; i < s->mbins; i++) s->binhead[i] = -1;
  for (int i = nall - 1; i >= 0; i--) {
    int ibin = coord2bin(s, &s->x[3*i]);
    if (ibin < 0 || ibin >= s->mbins) return fail("oracle: atom outside the bin grid (would be out-of-bounds in the reference)");
    s->bins[i] = s->binhead[ibin]; s->binhead[ibin] = i;
  }
  if (nlocal > s->maxlist) { s->maxlist = s->nmax; s->firstneigh = xrealloc(s->firstneigh, sizeof(long long) * s->maxlist); s->numneigh = xrealloc(s->numneigh, sizeof(int) * s->maxlist); }
  long long tot = 0;
  for (int i = 0; i < nlocal; i++) {
    int itype = s->type[i], n = 0;
    double xtmp = s->x[3*i], ytmp = s->x[3*i+1], ztmp = s->x[3*i+2];
    int ibin = coord2bin(s, &s->x[3*i]);
    s->firstneigh[i] = tot;
    for (int k = 0; k < s->nstencil; k++) {
      int b = ibin + s->stencil[k];
      if (b < 0 || b >= s->mbins) continue; /* cannot happen for owned atoms inside the box */
      for (int j = s->binhead[b]; j >= 0; j = s->bins[j]) {
        if (i == j) continue;
        double delx = xtmp - s->x[3*j], dely = ytmp - s->x[3*j+1], delz = ztmp - s->x[3*j+2];
        double rsq = delx * delx + dely * dely + delz * delz;
        if (rsq <= s->cutneighsq[itype * nt + s->type[j]]) {
          if (tot + n >= s->maxneigh) { s->maxneigh = s->maxneigh * 2 + 65536; s->neigh = xrealloc(s->neigh, sizeof(int) * s->maxneigh); s->halfkeep = xrealloc(s->halfkeep, s->maxneigh); }
          s->halfkeep[tot + n] = (unsigned char)half_keeps(s, i, j); /* derived lists are built together with the full list (neighbor.cpp:1489-1494) */
          s->neigh[tot + n++] = j;
        }
      }
    }
    s->numneigh[i] = n; tot += n;
    if (n > s->maxneighseen) s->maxneighseen = n;
  }
  return 0;
}

/* Neighbor::decide + check_distance, neighbor.cpp:1332-1410 (orthogonal, fixed box) */
static int neighbor_decide(osph_sph *s)
{
  for (int i = 0; i < s->nfix; i++)
    if (s->fix[i].kind == FIX_PHASE_CHANGE && s->ntimestep == s->fix[i].next_reneighbor) return 1; /* fix_check (:1334-1338) */
  s->ago++;
  if (s->ago >= s->delay && s->ago % s->every == 0) {
    if (s->check == 0) return 1;
    int flag = 0;
    for (int i = 0; i < s->nlocal; i++) {
      double delx = s->x[3*i] - s->xhold[3*i], dely = s->x[3*i+1] - s->xhold[3*i+1], delz = s->x[3*i+2] - s->xhold[3*i+2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq > s->triggersq) flag = 1;
    }
    if (flag && s->ago == (s->every > s->delay ? s->every : s->delay)) s->ndanger++;
    return flag;
  }
  return 0;
}

/* Which atom of a pair holds it in the reference's half list.
 * A deck with a full-list sub-style (sph/rhosum, sph/rhosum/multiphase, sph/colorgradient: pair_sph_rhosum.cpp:60-61) derives the half
 * list from the full one: half_from_full_newton, neigh_derive.cpp:117-135.
 * A deck without one builds it directly: half_bin_newton, neigh_half_bin.cpp:339-400 -- inside i's own bin the atoms behind i in the
 * bin's linked list (owned atoms in index order, then ghosts, which must lie "above and to the right"), and EVERY atom of the bins of the
 * upper half stencil (stencil_half_bin_{2d,3d}_newton, neigh_stencil.cpp:125-158: k > 0 || j > 0 || (j == 0 && i > 0)).
 * The pair sets are the same; who holds a pair matters where a style is not symmetric in (i, j): stale ghost fields (vest of a ghost at
 * setup is the border-time copy, examples/USER/sph/cavity_flow), the phase-change clamp, gamma of the list owner. */
static int half_keeps(osph_sph *s, int i, int j)
{
  if (s->half_bin) {
    int ix, iy, iz, jx, jy, jz;
    coord2bin3(s, &s->x[3*i], &ix, &iy, &iz); coord2bin3(s, &s->x[3*j], &jx, &jy, &jz);
    int dx = jx - ix, dy = jy - iy, dz = jz - iz;
    if (dx || dy || dz) return dz > 0 || (dz == 0 && (dy > 0 || (dy == 0 && dx > 0)));
  }
  if (j < s->nlocal) return !(i > j);
  if (s->x[3*j+2] < s->x[3*i+2]) return 0;
  if (s->x[3*j+2] == s->x[3*i+2]) {
    if (s->x[3*j+1] < s->x[3*i+1]) return 0;
    if (s->x[3*j+1] == s->x[3*i+1] && s->x[3*j] < s->x[3*i]) return 0;
  }
  return 1;
}

/* ======================================================================
   kernels: sph_kernel_quintic.cpp:17-73, sph_energy_equation.cpp:17-23
   ====================================================================== */
static double kernel_quintic(int dim, double r)
{
  const double norm = dim == 3 ? 0.0716197243913529 : 0.04195297663091802;
  const double q = 3.0 * r;
  if (q < 1.0) return norm * (pow(3 - q, 5) - 6 * pow(2 - q, 5) + 15 * pow(1 - q, 5));
  else if (q < 2.0) return norm * (pow(3 - q, 5) - 6 * pow(2 - q, 5));
  else if (q < 3.0) return norm * pow(3 - q, 5);
  return 0.0;
}
static double dw_quintic(int dim, double r)
{
  const double norm = dim == 3 ? 3.0 * 0.0716197243913529 : 3.0 * 0.04195297663091802;
  const double q = 3.0 * r;
  double wfd;
  if (q < 1) wfd = -50 * pow(q, 4) + 120 * pow(q, 3) - 120 * q;
  else if (q < 2) wfd = 25 * pow(q, 4) - 180 * pow(q, 3) + 450 * pow(q, 2) - 420 * q + 75;
  else if (q < 3.0) wfd = -5 * pow(q, 4) + 60 * pow(q, 3) - 270 * pow(q, 2) + 540 * q - 405;
  else wfd = 0.0;
  return norm * wfd;
}
static double sph_pressure(double B, double rho0, double gamma, double rbackground, double rho)
{ return B * (pow(rho / rho0, gamma) - rbackground); } /* pair_sph_taitwater_multiphase.cpp:289-292 */

/* ======================================================================
   pair styles
   ====================================================================== */
#define NB_LOOP_BEGIN(FULL)                                                              \
  for (int i = 0; i < nlocal; i++) {                                                     \
    int itype = s->type[i];                                                              \
    if (p->iskip[itype]) continue;                                                       \
    double xtmp = x[3*i], ytmp = x[3*i+1], ztmp = x[3*i+2];                              \
    const int *jlist = s->neigh + s->firstneigh[i]; int jnum = s->numneigh[i];                \
    const unsigned char *jkeep = s->halfkeep + s->firstneigh[i];
#define NB_FOR_J(FULL)                                                                   \
    for (int jj = 0; jj < jnum; jj++) {                                                  \
      int j = jlist[jj]; int jtype = s->type[j];                                         \
      if (!p->mapped[itype * nt + jtype]) continue;                                      \
      if (!(FULL) && !jkeep[jj]) continue;                                               \
      double delx = xtmp - x[3*j], dely = ytmp - x[3*j+1], delz = ztmp - x[3*j+2];       \
      double rsq = delx * delx + dely * dely + delz * delz;                              \
      if (rsq < p->cutsq[itype * nt + jtype]) {                                          \
        double h = p->cut[itype * nt + jtype];
#define NB_END }}}

/* PairSPHRhoSum::compute, pair_sph_rhosum.cpp:66-204 */
static void pair_rhosum(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x; double *rho = s->rho;
  if (p->nstep != 0 && (s->ntimestep % p->nstep) == 0) {
    for (int i = 0; i < nlocal; i++) {
      int itype = s->type[i]; if (p->iskip[itype]) continue;
      double h = p->cut[itype * nt + itype], wf;
      if (s->dim == 3) wf = 2.1541870227086614782 / (h * h * h); else wf = 1.5915494309189533576e0 / (h * h);
      rho[i] = s->mass[itype] * wf;
    }
    NB_LOOP_BEGIN(1) NB_FOR_J(1)
        double ih = 1.0 / h, ihsq = ih * ih, wf = 1.0 - rsq * ihsq;
        wf = wf * wf; wf = wf * wf;
        if (s->dim == 3) wf = 2.1541870227086614782e0 * wf * ihsq * ih; else wf = 1.5915494309189533576e0 * wf * ihsq;
        rho[i] += s->mass[jtype] * wf;
    NB_END
  }
  if (!s->inworld) comm_forward_rho(s); /* :203 (world mode: w_forward_rho after the style has run on every rank) */
}

/* PairSPHRhoSumMultiphase::compute, pair_sph_rhosum_multiphase.cpp:68-174.
 * Its forward comm is dead code (pack_comm vs pack_forward_comm). */
static void pair_rhosum_multiphase(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x; double *rho = s->rho;
  if (p->nstep != 0 && (s->ntimestep % p->nstep) == 0) {
    for (int i = 0; i < nlocal; i++) {
      int itype = s->type[i]; if (p->iskip[itype]) continue;
      double h = p->cut[itype * nt + itype];
      rho[i] = s->dim == 3 ? kernel_quintic(3, 0.0) / (h * h * h) : kernel_quintic(2, 0.0) / (h * h);
    }
    NB_LOOP_BEGIN(1) NB_FOR_J(1)
        double ih = 1.0 / h, r = sqrt(rsq) * ih, wf;
        if (s->dim == 3) wf = kernel_quintic(3, r) * ih * ih * ih; else wf = kernel_quintic(2, r) * ih * ih;
        rho[i] += wf;
      }}
      rho[i] *= s->rmass[i];
    }
  }
}

/* PairSPHColorGradient::compute, pair_sph_colorgradient.cpp:70-191 (forward comm is dead code) */
static void pair_colorgradient(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x; double *cg = s->cg;
  if (p->nstep != 0 && (s->ntimestep % p->nstep) == 0) {
    for (int i = 0; i < nlocal; i++) { if (p->iskip[s->type[i]]) continue; cg[3*i] = cg[3*i+1] = cg[3*i+2] = 0.0; }
    NB_LOOP_BEGIN(1)
      double sigmai = s->rho[i] / s->rmass[i];
      NB_FOR_J(1)
        double r = sqrt(rsq), eij[3];
        eij[0] = delx / r; eij[1] = dely / r; eij[2] = delz / r;
        double ih = 1.0 / h, wfd;
        if (s->dim == 3) { wfd = dw_quintic(3, r * ih); wfd = wfd * ih * ih * ih * ih; }
        else { wfd = dw_quintic(2, r * ih); wfd = wfd * ih * ih * ih; }
        double sigmaj = s->rho[j] / s->rmass[j], sigmaj2 = sigmaj * sigmaj;
        double dphi = -wfd * p->alpha[itype * nt + jtype] / sigmaj2 * sigmai;
        cg[3*i] += dphi * eij[0]; cg[3*i+1] += dphi * eij[1];
        if (s->dim == 3) cg[3*i+2] += dphi * eij[2];
    NB_END
  }
}

static double lucy_wfd(int dim, double h, double rsq)
{ /* pair_sph_taitwater.cpp:135-151 */
  double ih = 1.0 / h, ihsq = ih * ih, wfd = h - sqrt(rsq);
  if (dim == 3) wfd = -25.066903536973515383e0 * wfd * wfd * ihsq * ihsq * ihsq * ih;
  else wfd = -19.098593171027440292e0 * wfd * wfd * ihsq * ihsq * ihsq;
  return wfd;
}

/* PairSPHTaitwater::compute, pair_sph_taitwater.cpp:53-200 (morris=0)
 * PairSPHTaitwaterMorris::compute, pair_sph_taitwater_morris.cpp:52-200 (morris=1) */
static void pair_taitwater(osph_sph *s, opair *p, int morris)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *v = s->vest, *rho = s->rho, *mass = s->mass;
  double *f = s->f, *de = s->de, *drho = s->drho;
  NB_LOOP_BEGIN(0)
    double vxtmp = v[3*i], vytmp = v[3*i+1], vztmp = v[3*i+2], imass = mass[itype];
    double tmp = rho[i] / p->rho0[itype], fi = tmp * tmp * tmp;
    fi = p->B[itype] * (fi * fi * tmp - 1.0) / (rho[i] * rho[i]);
    NB_FOR_J(0)
      double jmass = mass[jtype], wfd = lucy_wfd(s->dim, h, rsq);
      tmp = rho[j] / p->rho0[jtype]; double fj = tmp * tmp * tmp;
      fj = p->B[jtype] * (fj * fj * tmp - 1.0) / (rho[j] * rho[j]);
      double velx = vxtmp - v[3*j], vely = vytmp - v[3*j+1], velz = vztmp - v[3*j+2];
      double delVdotDelR = delx * velx + dely * vely + delz * velz, fpair, deltaE, fvisc, fx, fy, fz;
      if (!morris) {
        if (delVdotDelR < 0.) {
          double mu = h * delVdotDelR / (rsq + 0.01 * h * h);
          fvisc = -p->viscosity[itype * nt + jtype] * (p->soundspeed[itype] + p->soundspeed[jtype]) * mu / (rho[i] + rho[j]);
        } else fvisc = 0.;
        fpair = -imass * jmass * (fi + fj + fvisc) * wfd;
        deltaE = -0.5 * fpair * delVdotDelR;
        fx = delx * fpair; fy = dely * fpair; fz = delz * fpair;
      } else {
        fvisc = 2 * p->viscosity[itype * nt + jtype] / (rho[i] * rho[j]);
        fvisc *= imass * jmass * wfd;
        fpair = -imass * jmass * (fi + fj) * wfd;
        deltaE = -0.5 * (fpair * delVdotDelR + fvisc * (velx * velx + vely * vely + velz * velz));
        fx = delx * fpair + velx * fvisc; fy = dely * fpair + vely * fvisc; fz = delz * fpair + velz * fvisc;
      }
      f[3*i] += fx; f[3*i+1] += fy; f[3*i+2] += fz;
      drho[i] += jmass * delVdotDelR * wfd;
      de[i] += deltaE;
      /* newton_pair on: always */
      f[3*j] -= fx; f[3*j+1] -= fy; f[3*j+2] -= fz;
      de[j] += deltaE;
      drho[j] += imass * delVdotDelR * wfd;
  NB_END
}

/* PairSPHIdealGas::compute, pair_sph_idealgas.cpp:48-175 */
static void pair_idealgas(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *v = s->vest, *rho = s->rho, *mass = s->mass, *e = s->e;
  double *f = s->f, *de = s->de, *drho = s->drho;
  NB_LOOP_BEGIN(0)
    double vxtmp = v[3*i], vytmp = v[3*i+1], vztmp = v[3*i+2], imass = mass[itype];
    double fi = 0.4 * e[i] / imass / rho[i];      /* ideal gas EOS: pressure / rho^2 (:94) */
    double ci = sqrt(0.4 * e[i] / imass);         /* speed of sound, gamma = 1.4 (:95) */
    NB_FOR_J(0)
      double jmass = mass[jtype], wfd = lucy_wfd(s->dim, h, rsq);
      double fj = 0.4 * e[j] / jmass / rho[j];
      double delVdotDelR = delx * (vxtmp - v[3*j]) + dely * (vytmp - v[3*j+1]) + delz * (vztmp - v[3*j+2]), fvisc;
      if (delVdotDelR < 0.) {
        double cj = sqrt(0.4 * e[j] / jmass);
        double mu = h * delVdotDelR / (rsq + 0.01 * h * h);
        fvisc = -p->viscosity[itype * nt + jtype] * (ci + cj) * mu / (rho[i] + rho[j]);
      } else fvisc = 0.;
      double fpair = -imass * jmass * (fi + fj + fvisc) * wfd, deltaE = -0.5 * fpair * delVdotDelR;
      f[3*i] += delx * fpair; f[3*i+1] += dely * fpair; f[3*i+2] += delz * fpair;
      drho[i] += jmass * delVdotDelR * wfd;
      de[i] += deltaE;
      f[3*j] -= delx * fpair; f[3*j+1] -= dely * fpair; f[3*j+2] -= delz * fpair;
      de[j] += deltaE;
      drho[j] += imass * delVdotDelR * wfd;
  NB_END
}

/* PairSPHLJ::LJEOS2, pair_sph_lj.cpp:303-333 (Ree 1980) */
static void ljeos2(double rho, double e, double cv, double *p, double *c)
{
  double T = e / cv, beta = 1.0 / T, beta_sqrt = sqrt(beta), x = rho * sqrt(beta_sqrt);
  double xsq = x * x, xpow3 = xsq * x, xpow4 = xsq * xsq;
  double diff_A_NkT = 3.629 + 7.264*x - beta*(3.492 - 18.698*x + 35.505*xsq - 31.816*xpow3 + 11.195*xpow4)
                    - beta_sqrt*(5.369 + 13.16*x + 18.525*xsq - 17.076*xpow3 + 9.32*xpow4)
                    + 10.4925*xsq + 11.46*xpow3 + 2.176*xpow4*xpow4*x;
  double d2A_dx2 = 7.264 + 20.985*x
                 + beta*(18.698 - 71.01*x + 95.448*xsq - 44.78*xpow3)
                 - beta_sqrt*(13.16 + 37.05*x - 51.228*xsq + 37.28*xpow3)
                 + 34.38*xsq + 19.584*xpow4*xpow4;
  *p = rho * T * (1.0 + diff_A_NkT * x);
  double csq = T * (1.0 + 2.0 * diff_A_NkT * x + d2A_dx2 * x * x);
  *c = csq > 0.0 ? sqrt(csq) : 0.0;
}
/* PairSPHLJ::compute, pair_sph_lj.cpp:48-182.  Note `fi += lrc` INSIDE the neighbor loop (:139): the long-range correction piles up
 * on fi, so the force of a pair depends on how many in-cutoff neighbors of i the half list visited before it -- the list order is
 * part of the result, which is why this oracle walks the list in the reference's own order. */
static void pair_lj(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *v = s->vest, *rho = s->rho, *mass = s->mass, *e = s->e, *cv = s->cv;
  double *f = s->f, *de = s->de, *drho = s->drho;
  NB_LOOP_BEGIN(0)
    double vxtmp = v[3*i], vytmp = v[3*i+1], vztmp = v[3*i+2], imass = mass[itype], fi, ci;
    ljeos2(rho[i], e[i], cv[i], &fi, &ci);
    fi /= (rho[i] * rho[i]);
    NB_FOR_J(0)
      double jmass = mass[jtype], ih = 1.0 / h, ihsq = ih * ih, ihcub = ihsq * ih, wfd = h - sqrt(rsq), fj, cj;
      if (s->dim == 3) wfd = -25.066903536973515383e0 * wfd * wfd * ihsq * ihsq * ihsq * ih;
      else wfd = -19.098593171027440292e0 * wfd * wfd * ihsq * ihsq * ihsq;
      ljeos2(rho[j], e[j], cv[j], &fj, &cj);
      fj /= (rho[j] * rho[j]);
      double lrc = -11.1701 * (ihcub * ihcub * ihcub - 1.5 * ihcub);
      fi += lrc; fj += lrc;
      double delVdotDelR = delx * (vxtmp - v[3*j]) + dely * (vytmp - v[3*j+1]) + delz * (vztmp - v[3*j+2]), fvisc;
      if (delVdotDelR < 0.) {
        double mu = h * delVdotDelR / (rsq + 0.01 * h * h);
        fvisc = -p->viscosity[itype * nt + jtype] * (ci + cj) * mu / (rho[i] + rho[j]);
      } else fvisc = 0.;
      double fpair = -imass * jmass * (fi + fj + fvisc) * wfd, deltaE = -0.5 * fpair * delVdotDelR;
      f[3*i] += delx * fpair; f[3*i+1] += dely * fpair; f[3*i+2] += delz * fpair;
      drho[i] += jmass * delVdotDelR * wfd;
      de[i] += deltaE;
      f[3*j] -= delx * fpair; f[3*j+1] -= dely * fpair; f[3*j+2] -= delz * fpair;
      de[j] += deltaE;
      drho[j] += imass * delVdotDelR * wfd;
  NB_END
}

/* PairSPHTaitwaterMultiphase::compute, pair_sph_taitwater_multiphase.cpp:55-186 */
static void pair_taitwater_multiphase(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *v = s->vest, *rho = s->rho, *rmass = s->rmass;
  double *f = s->f;
  NB_LOOP_BEGIN(0)
    double vxtmp = v[3*i], vytmp = v[3*i+1], vztmp = v[3*i+2], imass = rmass[i];
    double pi = sph_pressure(p->B[itype], p->rho0[itype], p->gamma[itype], p->rbackground[itype], rho[i]);
    double Vi = imass / rho[i], Vi2 = Vi * Vi;
    NB_FOR_J(0)
      double jmass = rmass[j], ih = 1.0 / h, wfd;
      if (s->dim == 3) { wfd = dw_quintic(3, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih * ih / sqrt(rsq); }
      else { wfd = dw_quintic(2, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih / sqrt(rsq); }
      double Vj = jmass / rho[j], Vj2 = Vj * Vj;
      double pj = sph_pressure(p->B[jtype], p->rho0[jtype], p->gamma[itype] /* sic, :148 */, p->rbackground[jtype], rho[j]);
      double pij_wave = (rho[j] * pi + rho[i] * pj) / (rho[i] + rho[j]);
      double velx = vxtmp - v[3*j], vely = vytmp - v[3*j+1], velz = vztmp - v[3*j+2];
      double fvisc = (Vi2 + Vj2) * p->viscosity[itype * nt + jtype] * wfd;
      double fpair = -(Vi2 + Vj2) * pij_wave * wfd;
      f[3*i] += delx * fpair + velx * fvisc; f[3*i+1] += dely * fpair + vely * fvisc; f[3*i+2] += delz * fpair + velz * fvisc;
      f[3*j] -= delx * fpair + velx * fvisc; f[3*j+1] -= dely * fpair + vely * fvisc; f[3*j+2] -= delz * fpair + velz * fvisc;
  NB_END
}

/* PairSPHSurfaceTension::compute, pair_sph_surfacetension.cpp:50-192 */
static void surface_force(int dim, const double *c, double abscg, const double *eij, double *S)
{
  S[0] = S[1] = S[2] = 0.0;
  if (!(abscg > EPSILON)) return;
  if (dim == 2) {
    S[0] = (eij[0] * ((c[1] * c[1] + c[0] * c[0]) / 2 - c[0] * c[0]) - c[0] * eij[1] * c[1]) / abscg;
    S[1] = (eij[1] * ((c[1] * c[1] + c[0] * c[0]) / 2 - c[1] * c[1]) - eij[0] * c[0] * c[1]) / abscg;
  } else {
    S[0] = (eij[0] * (0.3333333333333333 * c[2] * c[2] + 0.3333333333333333 * c[1] * c[1] - 0.6666666666666666 * c[0] * c[0])
            - 1.0 * c[0] * eij[2] * c[2] - 1.0 * c[0] * eij[1] * c[1]) / abscg;
    S[1] = (eij[1] * (0.3333333333333333 * c[2] * c[2] - 0.6666666666666666 * c[1] * c[1] + 0.3333333333333333 * c[0] * c[0])
            - 1.0 * c[1] * eij[2] * c[2] - 1.0 * eij[0] * c[0] * c[1]) / abscg;
    S[2] = (eij[2] * (-0.6666666666666666 * c[2] * c[2] + 0.3333333333333333 * c[1] * c[1] + 0.3333333333333333 * c[0] * c[0])
            - 1.0 * eij[1] * c[1] * c[2] - 1.0 * eij[0] * c[0] * c[2]) / abscg;
  }
}
static void pair_surfacetension(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1, ndim = s->dim; const double *x = s->x, *cg = s->cg, *rho = s->rho, *rmass = s->rmass;
  double *f = s->f;
  NB_LOOP_BEGIN(0)
    double imass = rmass[i], abscgi;
    if (ndim == 3) abscgi = sqrt(cg[3*i] * cg[3*i] + cg[3*i+1] * cg[3*i+1] + cg[3*i+2] * cg[3*i+2]);
    else abscgi = sqrt(cg[3*i] * cg[3*i] + cg[3*i+1] * cg[3*i+1]);
    NB_FOR_J(0)
      double jmass = rmass[j], ih = 1.0 / h, wfd, eij[3] = {0, 0, 0}, Si[3], Sj[3], abscgj;
      if (ndim == 3) { wfd = dw_quintic(3, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih * ih; }
      else { wfd = dw_quintic(2, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih; }
      eij[0] = delx / sqrt(rsq); eij[1] = dely / sqrt(rsq); if (ndim == 3) eij[2] = delz / sqrt(rsq);
      if (ndim == 2) abscgj = sqrt(cg[3*j] * cg[3*j] + cg[3*j+1] * cg[3*j+1]);
      else abscgj = sqrt(cg[3*j] * cg[3*j] + cg[3*j+1] * cg[3*j+1] + cg[3*j+2] * cg[3*j+2]);
      surface_force(ndim, &cg[3*i], abscgi, eij, Si);
      surface_force(ndim, &cg[3*j], abscgj, eij, Sj);
      const double Vi = imass / rho[i], Vj = jmass / rho[j];
      for (int d = 0; d < ndim; d++) {
        f[3*i+d] += (Si[d] * Vi * Vi + Sj[d] * Vj * Vj) * wfd;
        f[3*j+d] -= (Si[d] * Vi * Vi + Sj[d] * Vj * Vj) * wfd;
      }
  NB_END
}

/* PairSPHHeatConduction::compute, pair_sph_heatconduction.cpp:47-134 */
static void pair_heatconduction(osph_sph *s, opair *p)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *e = s->e, *rho = s->rho, *mass = s->mass; double *de = s->de;
  NB_LOOP_BEGIN(0)
    double imass = mass[itype];
    NB_FOR_J(0)
      double jmass = mass[jtype], wfd = lucy_wfd(s->dim, h, rsq), D = p->alpha[itype * nt + jtype];
      double deltaE = 2.0 * imass * jmass / (imass + jmass);
      deltaE *= (rho[i] + rho[j]) / (rho[i] * rho[j]);
      deltaE *= D * (e[i] - e[j]) * wfd;
      de[i] += deltaE; de[j] -= deltaE;
  NB_END
}

/* PairSPHHeatConductionMultiPhase::compute (..._multiphase.cpp:49-129, phasechange=0)
 * PairSPHHeatConductionPhaseChange::compute (..._phasechange.cpp:52-141, phasechange=1) */
static void pair_heatconduction_multiphase(osph_sph *s, opair *p, int phasechange)
{
  int nlocal = s->nlocal, nt = s->ntypes + 1; const double *x = s->x, *e = s->e, *rho = s->rho, *rmass = s->rmass, *cv = s->cv; double *de = s->de;
  NB_LOOP_BEGIN(0)
    double imass = rmass[i];
    NB_FOR_J(0)
      double ih = 1.0 / h, wfd;
      if (s->dim == 3) { wfd = dw_quintic(3, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih * ih / sqrt(rsq); }
      else { wfd = dw_quintic(2, sqrt(rsq) * ih); wfd = wfd * ih * ih * ih / sqrt(rsq); }
      double jmass = rmass[j], D = p->alpha[itype * nt + jtype];
      double Ti = e[i] / cv[i], Tj = e[j] / cv[j];
      if (phasechange) {
        int ff = p->fixflag ? p->fixflag[itype * nt + jtype] : 0; double tcij = p->tc ? p->tc[itype * nt + jtype] : 0.0;
        if ((ff == itype) && (Ti < Tj)) Ti = tcij;
        if ((ff == jtype) && (Tj < Ti)) Tj = tcij;
      }
      double deltaE = 2.0 * D * (Ti - Tj) * wfd / (rho[i] * rho[j]);
      de[i] += deltaE * jmass; de[j] -= deltaE * imass;
  NB_END
}

static int pair_compute_slot(osph_sph *s, int k)
{
  opair *p = &s->pair[k];
  switch (p->style) {
  case B200_PAIR_RHOSUM: pair_rhosum(s, p); break;
  case B200_PAIR_RHOSUM_MULTIPHASE: pair_rhosum_multiphase(s, p); break;
  case B200_PAIR_TAITWATER: pair_taitwater(s, p, 0); break;
  case B200_PAIR_TAITWATER_MORRIS: pair_taitwater(s, p, 1); break;
  case B200_PAIR_TAITWATER_MULTIPHASE: pair_taitwater_multiphase(s, p); break;
  case B200_PAIR_COLORGRADIENT: pair_colorgradient(s, p); break;
  case B200_PAIR_SURFACETENSION: pair_surfacetension(s, p); break;
  case B200_PAIR_HEATCONDUCTION: pair_heatconduction(s, p); break;
  case B200_PAIR_HEATCONDUCTION_MULTIPHASE: pair_heatconduction_multiphase(s, p, 0); break;
  case B200_PAIR_HEATCONDUCTION_PHASECHANGE: pair_heatconduction_multiphase(s, p, 1); break;
  case B200_PAIR_IDEALGAS: pair_idealgas(s, p); break;
  case B200_PAIR_LJ: pair_lj(s, p); break;
  default: return fail("unknown pair style");
  }
  return 0;
}

/* ======================================================================
   fixes
   ====================================================================== */

/* RanPark::uniform, random_park.cpp:40-47 */
static double ranpark(int *seed)
{
  int k = *seed / 127773;
  *seed = 16807 * (*seed - k * 127773) - 2836 * k;
  if (*seed < 0) *seed += 2147483647;
  return (1.0 / 2147483647) * *seed;
}

/* FixMeso::initial_integrate (fix_meso.cpp:91-140), FixMesoStationary (:71-92) */
static void fix_initial_integrate(osph_sph *s, ofix *fx)
{
  double dtv = s->dt, dtf = 0.5 * s->dt * s->ftm2v;
  for (int i = 0; i < s->nlocal; i++) {
    if (!(s->mask[i] & fx->groupbit)) continue;
    s->e[i] += dtf * s->de[i];
    s->rho[i] += dtf * s->drho[i];
    if (fx->kind == FIX_MESO) {
      double dtfm = s->multiphase ? dtf / s->rmass[i] : dtf / s->mass[s->type[i]];
      for (int d = 0; d < 3; d++) {
        s->vest[3*i+d] = s->v[3*i+d] + 2.0 * dtfm * s->f[3*i+d];
        s->v[3*i+d] += dtfm * s->f[3*i+d];
        s->x[3*i+d] += dtv * s->v[3*i+d];
      }
    }
  }
}
/* FixMeso::final_integrate (fix_meso.cpp:144-180), FixMesoStationary (:96-112) */
static void fix_final_integrate(osph_sph *s, ofix *fx)
{
  double dtf = 0.5 * s->dt * s->ftm2v;
  for (int i = 0; i < s->nlocal; i++) {
    if (!(s->mask[i] & fx->groupbit)) continue;
    if (fx->kind == FIX_MESO) {
      double dtfm = s->multiphase ? dtf / s->rmass[i] : dtf / s->mass[s->type[i]];
      for (int d = 0; d < 3; d++) s->v[3*i+d] += dtfm * s->f[3*i+d];
    }
    s->e[i] += dtf * s->de[i];
    s->rho[i] += dtf * s->drho[i];
  }
}
/* FixGravity::post_force, fix_gravity.cpp:244-295 */
static void fix_gravity(osph_sph *s, ofix *fx)
{
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & fx->groupbit) {
      double massone = s->multiphase ? s->rmass[i] : s->mass[s->type[i]];
      for (int d = 0; d < 3; d++) s->f[3*i+d] += massone * fx->acc[d];
    }
}

/* ---- variable formulas: Variable::evaluate restated (variable.cpp:1480-1735), evaluated directly for one atom with an argument
 * stack and an operator stack: an operator of precedence p first pops and applies every stacked operator of precedence >= p
 * (:1641), so all binary operators, `^` included, associate to the left; precedences :99-107.  Atom vectors as in
 * Variable::is_atom_vector / atom_vector (:3480-3560): id mass type x y z vx vy vz fx fy fz; thermo keywords step, dt. ---- */
enum { V_DONE, V_ADD, V_SUB, V_MUL, V_DIV, V_CARAT, V_MOD, V_UNARY, V_NOT, V_EQ, V_NE, V_LT, V_LE, V_GT, V_GE, V_AND, V_OR };
static const int v_prec[] = {0, 5, 5, 6, 6, 7, 6, 8, 8, 3, 3, 4, 4, 4, 4, 2, 1};
typedef struct { osph_sph *s; int i; const char *p; int bad; } vctx;
static double v_formula(vctx *c, const char *stops);
static double v_atomword(vctx *c, const char *w)
{
  osph_sph *s = c->s; int i = c->i;
  if (!strcmp(w, "id")) return s->tag[i];
  if (!strcmp(w, "mass")) return s->multiphase ? s->rmass[i] : s->mass[s->type[i]];
  if (!strcmp(w, "type")) return s->type[i];
  if (!strcmp(w, "x")) return s->x[3*i]; if (!strcmp(w, "y")) return s->x[3*i+1]; if (!strcmp(w, "z")) return s->x[3*i+2];
  if (!strcmp(w, "vx")) return s->v[3*i]; if (!strcmp(w, "vy")) return s->v[3*i+1]; if (!strcmp(w, "vz")) return s->v[3*i+2];
  if (!strcmp(w, "fx")) return s->f[3*i]; if (!strcmp(w, "fy")) return s->f[3*i+1]; if (!strcmp(w, "fz")) return s->f[3*i+2];
  if (!strcmp(w, "step")) return (double)s->ntimestep;
  if (!strcmp(w, "dt")) return s->dt;
  if (!strcmp(w, "PI")) return 3.14159265358979323846;
  c->bad = 1; return 0.0;
}
static double v_apply(int op, double v1, double v2)
{
  switch (op) {
  case V_ADD: return v1 + v2; case V_SUB: return v1 - v2; case V_MUL: return v1 * v2; case V_DIV: return v1 / v2;
  case V_MOD: return fmod(v1, v2); case V_CARAT: return pow(v1, v2); case V_UNARY: return -v2; case V_NOT: return v2 == 0.0 ? 1.0 : 0.0;
  case V_EQ: return v1 == v2 ? 1.0 : 0.0; case V_NE: return v1 != v2 ? 1.0 : 0.0; case V_LT: return v1 < v2 ? 1.0 : 0.0;
  case V_LE: return v1 <= v2 ? 1.0 : 0.0; case V_GT: return v1 > v2 ? 1.0 : 0.0; case V_GE: return v1 >= v2 ? 1.0 : 0.0;
  case V_AND: return (v1 != 0.0 && v2 != 0.0) ? 1.0 : 0.0; default: return (v1 != 0.0 || v2 != 0.0) ? 1.0 : 0.0;
  }
}
static double v_formula(vctx *c, const char *stops)
{
  double arg[64]; int op[64], narg = 0, nop = 0, expect_arg = 1;
  for (;;) {
    while (*c->p == ' ' || *c->p == '\t') c->p++;
    char ch = *c->p;
    if (expect_arg && ch == '(') { c->p++; arg[narg++] = v_formula(c, ")"); if (*c->p != ')') c->bad = 1; else c->p++; expect_arg = 0; continue; }
    if (expect_arg && (isdigit((unsigned char)ch) || ch == '.')) {
      const char *q = c->p;
      while (isdigit((unsigned char)*q) || *q == '.' || *q == 'e' || *q == 'E' || ((*q == '-' || *q == '+') && q > c->p && (q[-1] == 'e' || q[-1] == 'E'))) q++;
      char buf[64]; int n = (int)(q - c->p); if (n > 63) n = 63; memcpy(buf, c->p, n); buf[n] = 0;
      arg[narg++] = atof(buf); c->p = q; expect_arg = 0; continue;
    }
    if (expect_arg && (isalpha((unsigned char)ch) || ch == '_')) {
      char w[32]; int n = 0;
      while ((isalnum((unsigned char)*c->p) || *c->p == '_') && n < 31) w[n++] = *c->p++;
      w[n] = 0;
      if (*c->p == '(') {                        /* math functions, variable.cpp:2560-2900 */
        c->p++;
        double a = v_formula(c, ",)"), b = 0.0;
        if (*c->p == ',') { c->p++; b = v_formula(c, ")"); }
        if (*c->p != ')') c->bad = 1; else c->p++;
        double r = 0.0;
        if (!strcmp(w, "sqrt")) r = sqrt(a); else if (!strcmp(w, "exp")) r = exp(a); else if (!strcmp(w, "ln")) r = log(a);
        else if (!strcmp(w, "log")) r = log10(a); else if (!strcmp(w, "abs")) r = fabs(a); else if (!strcmp(w, "sin")) r = sin(a);
        else if (!strcmp(w, "cos")) r = cos(a); else if (!strcmp(w, "tan")) r = tan(a); else if (!strcmp(w, "asin")) r = asin(a);
        else if (!strcmp(w, "acos")) r = acos(a); else if (!strcmp(w, "atan")) r = atan(a); else if (!strcmp(w, "atan2")) r = atan2(a, b);
        else if (!strcmp(w, "ceil")) r = ceil(a); else if (!strcmp(w, "floor")) r = floor(a);
        else if (!strcmp(w, "round")) r = (a < 0.0) ? ceil(a - 0.5) : floor(a + 0.5); else c->bad = 1;
        arg[narg++] = r;
      } else arg[narg++] = v_atomword(c, w);
      expect_arg = 0; continue;
    }
    int o = -1, len = 1;
    if (ch == 0 || strchr(stops, ch)) o = V_DONE;
    else if (ch == '-' && expect_arg) o = V_UNARY; else if (ch == '!' && c->p[1] != '=' && expect_arg) o = V_NOT;
    else if (expect_arg) { c->bad = 1; return 0.0; }
    else if (ch == '+') o = V_ADD; else if (ch == '-') o = V_SUB; else if (ch == '*') o = V_MUL; else if (ch == '/') o = V_DIV;
    else if (ch == '^') o = V_CARAT; else if (ch == '%') o = V_MOD;
    else if (ch == '=' && c->p[1] == '=') { o = V_EQ; len = 2; } else if (ch == '!' && c->p[1] == '=') { o = V_NE; len = 2; }
    else if (ch == '<' && c->p[1] == '=') { o = V_LE; len = 2; } else if (ch == '<') o = V_LT;
    else if (ch == '>' && c->p[1] == '=') { o = V_GE; len = 2; } else if (ch == '>') o = V_GT;
    else if (ch == '&' && c->p[1] == '&') { o = V_AND; len = 2; } else if (ch == '|' && c->p[1] == '|') { o = V_OR; len = 2; }
    if (o < 0 || narg >= 60 || nop >= 60) { c->bad = 1; return 0.0; }
    if (o != V_UNARY && o != V_NOT)
      while (nop && v_prec[op[nop - 1]] >= v_prec[o]) {
        int prev = op[--nop]; double v2 = arg[--narg], v1 = 0.0;
        if (prev != V_UNARY && prev != V_NOT) v1 = arg[--narg];
        arg[narg++] = v_apply(prev, v1, v2);
      }
    if (o == V_DONE) return narg == 1 ? arg[0] : (c->bad = 1, 0.0);
    op[nop++] = o; c->p += len; expect_arg = 1;
  }
}
static double var_eval(osph_sph *s, const char *text, int i)
{
  vctx c = {s, i, text, 0};
  double v = v_formula(&c, "");
  if (c.bad || *c.p) { fprintf(stderr, "sph_oracle: cannot evaluate variable formula '%s'\n", text); abort(); }
  return v;
}
int osph_formula_check(const char *formula, const double atom[12], int type, int id, double step, double dt, double *value)
{
  osph_sph *s = calloc(1, sizeof *s);
  double x[3], v[3], f[3], rm[1]; int ty[1] = {type}, tg[1] = {id};
  if (atom) { memcpy(x, atom, sizeof x); memcpy(v, atom + 3, sizeof v); memcpy(f, atom + 6, sizeof f); rm[0] = atom[9]; }
  else { memset(x, 0, sizeof x); memset(v, 0, sizeof v); memset(f, 0, sizeof f); rm[0] = 1.0; }
  s->multiphase = 1; s->x = x; s->v = v; s->f = f; s->rmass = rm; s->type = ty; s->tag = tg; s->ntimestep = (long long)step; s->dt = dt;
  vctx c = {s, 0, formula, 0};
  double r = v_formula(&c, "");
  int bad = c.bad || *c.p;
  free(s);
  if (bad) return fail("Invalid syntax in variable formula");
  if (value) *value = r;
  return 0;
}
/* FixAddForce::post_force, fix_addforce.cpp:243-330: the variable components are evaluated for the whole group first
 * (Variable::compute_atom), then added */
static void fix_addforce(osph_sph *s, ofix *fx)
{
  double *add = malloc(sizeof(double) * 3 * (s->nlocal + 1));
  for (int i = 0; i < s->nlocal; i++)
    for (int d = 0; d < 3; d++)
      add[3*i+d] = !(s->mask[i] & fx->groupbit) ? 0.0 : (fx->formula[d] ? var_eval(s, fx->formula[d], i) : fx->acc[d]);
  for (int i = 0; i < 3 * s->nlocal; i++) s->f[i] += add[i];
  free(add);
}

/* FixSetMeso::post_force, constant value (fix_setmeso.cpp:180-236); Region::match with side in
 * (region.cpp:127-131, region_block.cpp:114-119, region_sphere.cpp:96-105) */
static void fix_setmeso(osph_sph *s, ofix *fx)
{
  for (int i = 0; i < s->nlocal; i++) {
    if (!(s->mask[i] & fx->groupbit)) continue;
    if (fx->region_kind) {
      const double *x = &s->x[3*i], *r = fx->region; int in;
      if (fx->region_kind == 1) in = x[0] >= r[0] && x[0] <= r[1] && x[1] >= r[2] && x[1] <= r[3] && x[2] >= r[4] && x[2] <= r[5];
      else { double delx = x[0] - r[0], dely = x[1] - r[1], delz = x[2] - r[2]; in = sqrt(delx * delx + dely * dely + delz * delz) <= r[3]; }
      if (fx->match_inside && !in) continue;
      if (!fx->match_inside && in) continue;
    }
    const double value = fx->formula[0] ? var_eval(s, fx->formula[0], i) : fx->value;     /* varflag ATOM / EQUAL, :238-262 */
    if (fx->which == 0) s->rho[i] = value;
    else if (fx->which == 1) s->e[i] = value;
    else s->e[i] = s->cv[i] * value;      /* sph_t2energy */
  }
}
/* FixEnforce2D::post_force, fix_enforce2d.cpp:77-89 */
static void fix_enforce2d(osph_sph *s, ofix *fx)
{ for (int i = 0; i < s->nlocal; i++) if (s->mask[i] & fx->groupbit) { s->v[3*i+2] = 0.0; s->f[3*i+2] = 0.0; } }
/* FixDtReset::end_of_step, fix_dt_reset.cpp:131-186 */
static void fix_dt_reset(osph_sph *s, ofix *fx)
{
  const double BIGDT = 1.0e20;
  double dtmin = BIGDT;
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & fx->groupbit) {
      const double *v = &s->v[3*i], *f = &s->f[3*i];
      double massinv = s->multiphase ? 1.0 / s->rmass[i] : 1.0 / s->mass[s->type[i]];
      double vsq = v[0]*v[0] + v[1]*v[1] + v[2]*v[2], fsq = f[0]*f[0] + f[1]*f[1] + f[2]*f[2];
      double dtv = BIGDT, dtf = BIGDT;
      if (vsq > 0.0) dtv = fx->xmax / sqrt(vsq);
      if (fsq > 0.0) dtf = sqrt(2.0 * fx->xmax / (s->ftm2v * sqrt(fsq) * massinv));
      double dt = dtv < dtf ? dtv : dtf, dtsq = dt * dt;
      double delx = dt*v[0] + 0.5*dtsq*massinv*f[0] * s->ftm2v, dely = dt*v[1] + 0.5*dtsq*massinv*f[1] * s->ftm2v,
             delz = dt*v[2] + 0.5*dtsq*massinv*f[2] * s->ftm2v;
      double delr = sqrt(delx*delx + dely*dely + delz*delz);
      if (delr > fx->xmax) dt *= fx->xmax / delr;
      if (dt < dtmin) dtmin = dt;
    }
  double dt = dtmin;
  if (fx->minbound && dt < fx->tmin) dt = fx->tmin;
  if (fx->maxbound && dt > fx->tmax) dt = fx->tmax;
  if (dt == s->dt) return;                                   /* fix_dt_reset.cpp:175 */
  s->laststep = s->ntimestep;
  s->atime += (s->ntimestep - s->atimestep) * s->dt;         /* Update::update_time, update.cpp:480-484 */
  s->atimestep = s->ntimestep;
  s->dt = dt;
}
/* FixSetMesodE::post_force, constant value, fix_setmesode.cpp:171-199 */
static void fix_setmesode(osph_sph *s, ofix *fx)
{
  for (int i = 0; i < s->nlocal; i++) {
    if (!(s->mask[i] & fx->groupbit)) continue;
    if (fx->region_kind) {
      const double *x = &s->x[3*i], *r = fx->region; int in;
      if (fx->region_kind == 1) in = x[0] >= r[0] && x[0] <= r[1] && x[1] >= r[2] && x[1] <= r[3] && x[2] >= r[4] && x[2] <= r[5];
      else { double dx = x[0] - r[0], dy = x[1] - r[1], dz = x[2] - r[2]; in = sqrt(dx * dx + dy * dy + dz * dz) <= r[3]; }
      if (!in) continue;
    }
    s->de[i] = fx->value;
  }
}
/* FixSetForce::post_force, constant values, fix_setforce.cpp:241-251 */
static void fix_setforce(osph_sph *s, ofix *fx)
{
  for (int i = 0; i < s->nlocal; i++)
    if (s->mask[i] & fx->groupbit)
      for (int d = 0; d < 3; d++) if (fx->fset[d]) s->f[3*i+d] = fx->fvalue[d];
}

/* ---- fix phase_change, fix_phase_change.cpp:167-352 ---- */
static int pc_isfromphasearound(osph_sph *s, ofix *fx, int i)
{ /* :540-563 */
  double cutoff2 = fx->pc.cutoff * fx->pc.cutoff;
  const int *jlist = s->neigh + s->firstneigh[i];
  for (int jj = 0; jj < s->numneigh[i]; jj++) {
    int j = jlist[jj];
    if (s->type[j] == fx->pc.from_type) {
      double delx = s->x[3*i] - s->x[3*j], dely = s->x[3*i+1] - s->x[3*j+1], delz = s->x[3*i+2] - s->x[3*j+2];
      if (delx * delx + dely * dely + delz * delz <= cutoff2) return 1;
    }
  }
  return 0;
}
static void pc_create_newpos(osph_sph *s, ofix *fx, const double *xone, const double *cgone, double delta, double *coord)
{ /* :473-513 */
  double eij[3];
  if (s->dim == 3) {
    double b1[3] = {-cgone[1], cgone[0], 0};
    double b1abs = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    if (b1abs > CG_SMALL) { b1[0] = b1[0] / b1abs; b1[1] = b1[1] / b1abs; b1[2] = b1[2] / b1abs; }
    double b2[3];
    b2[0] = -cgone[0] * cgone[1] * cgone[2] / (pow(cgone[1], 2) + pow(cgone[0], 2));
    b2[1] = -cgone[2] * pow(cgone[1], 2) / (pow(cgone[1], 2) + pow(cgone[0], 2));
    b2[2] = cgone[1];
    double b2abs = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
    if (b1abs > CG_SMALL) { b2[0] = b2[0] / b2abs; b2[1] = b2[1] / b2abs; b2[2] = b2[2] / b2abs; }
    double atmp = ranpark(&fx->seed) - 0.5;
    double btmp = ranpark(&fx->seed) - 0.5;
    eij[0] = atmp * b1[0] + btmp * b2[0]; eij[1] = atmp * b1[1] + btmp * b2[1]; eij[2] = atmp * b1[2] + btmp * b2[2];
  } else {
    double atmp = ranpark(&fx->seed);
    if (atmp > 0.5) atmp = 1; else atmp = -1;
    eij[0] = -atmp * cgone[1]; eij[1] = atmp * cgone[0]; eij[2] = 0.0;
  }
  double eijabs = sqrt(eij[0] * eij[0] + eij[1] * eij[1] + eij[2] * eij[2]);
  coord[0] = xone[0] + eij[0] * delta / eijabs;
  coord[1] = xone[1] + eij[1] * delta / eijabs;
  coord[2] = xone[2] + eij[2] * delta / eijabs;
}
static void pc_create_newpos_simple(ofix *fx, const double *xone, double delta, double *coord)
{ /* :467-471 */
  coord[0] = xone[0] + (ranpark(&fx->seed) - 0.5) * delta;
  coord[1] = xone[1] + (ranpark(&fx->seed) - 0.5) * delta;
  coord[2] = xone[2] + (ranpark(&fx->seed) - 0.5) * delta;
}
/* insert_one_atom (:425-465) + AtomVecMesoMultiPhase::create_atom (atom_vec_meso_multiphase.cpp:963-994).
 * The new atom is written at index nlocal, i.e. over the first ghost slot -- as in the reference. */
static int pc_insert_one_atom(osph_sph *s, ofix *fx, const double *coord)
{
  int flag = 0;
  if (coord[0] >= s->sublo[0] && coord[0] < s->subhi[0] && coord[1] >= s->sublo[1] && coord[1] < s->subhi[1] &&
      coord[2] >= s->sublo[2] && coord[2] < s->subhi[2]) flag = 1;
  else if (s->dim == 3 && coord[2] >= s->boxhi[2] && coord[0] >= s->sublo[0] && coord[0] < s->subhi[0] &&
           coord[1] >= s->sublo[1] && coord[1] < s->subhi[1]) flag = 1;
  else if (s->dim == 2 && coord[1] >= s->boxhi[1] && coord[0] >= s->sublo[0] && coord[0] < s->subhi[0]) flag = 1;
  if (!flag) return 0;
  grow(s, s->nlocal + 1 + s->nghost);
  int m = s->nlocal;
  s->tag[m] = 0; s->type[m] = fx->pc.to_type; s->img[m] = 13;
  for (int d = 0; d < 3; d++) { s->x[3*m+d] = coord[d]; s->v[3*m+d] = 0.0; s->vest[3*m+d] = 0.0; s->cg[3*m+d] = 0.0; }
  s->rho[m] = 0.0; s->rmass[m] = 0.0; s->e[m] = 0.0; s->cv[m] = 1.0; s->de[m] = 0.0; s->drho[m] = 0.0;
  s->mask[m] = 1 | fx->groupbit;
  s->nlocal++;
  return 1;
}

/* FixPhaseChange::pre_exchange up to the reverse communication of dmass (:167-323), on one rank: 0 = not due this step, 1 = done */
static int pc_local(osph_sph *s, ofix *fx)
{
  const osph_phase_change_desc *pc = &fx->pc;
  if (fx->next_reneighbor != s->ntimestep) return 0;
  int nins = 0, nlocal = s->nlocal, nall = s->nlocal + s->nghost;
  double *dmass = s->drho; /* :193 */
  for (int i = 0; i < nall; i++) dmass[i] = 0.0;
  for (int i = 0; i < nlocal; i++) {
    double Ti = s->e[i] / s->cv[i];
    int isphasechange;
    if ((Ti < pc->Tc) || (s->type[i] != pc->to_type)) isphasechange = 0;
    else if (pc->energy_chance_flag) {
      double threshold = (s->e[i] - s->cv[i] * pc->Tc) / pc->Hwv * s->dt * pc->phase_change_rate;
      isphasechange = (ranpark(&fx->seed) < threshold) && pc_isfromphasearound(s, fx, i);
    } else isphasechange = (ranpark(&fx->seed) < pc->change_chance) && (Ti > pc->Tt) && pc_isfromphasearound(s, fx, i);
    if (!isphasechange) continue;
    double coord[3]; int ok; double delta = pc->dr; int natempt = 0;
    do { pc_create_newpos(s, fx, &s->x[3*i], &s->cg[3*i], delta, coord); ok = pc_insert_one_atom(s, fx, coord); delta = 0.75 * delta; natempt++; } while (!ok && natempt < pc->maxattempt);
    if (!ok) {
      delta = pc->dr; natempt = 0;
      do { pc_create_newpos_simple(fx, &s->x[3*i], delta, coord); ok = pc_insert_one_atom(s, fx, coord); delta = 0.75 * delta; natempt++; } while (!ok && natempt < pc->maxattempt);
    }
    if (!ok) continue;
    nins++;
    double xtmp = s->x[3*i], ytmp = s->x[3*i+1], ztmp = s->x[3*i+2];
    int jnum = s->numneigh[i]; const int *jlist = s->neigh + s->firstneigh[i];
    double wtotal = 0.0;
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      if ((s->type[j] == pc->from_type) && (s->rmass[j] > 0.5 * pc->to_mass)) {
        double delx = xtmp - s->x[3*j], dely = ytmp - s->x[3*j+1], delz = ztmp - s->x[3*j+2];
        double rsq = delx * delx + dely * dely + delz * delz;
        wtotal += kernel_quintic(s->dim, sqrt(rsq) * pc->cutoff);
      }
    }
    double dmom[3] = {0, 0, 0}, dmomest[3] = {0, 0, 0}, denergy = 0.0;
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      if ((s->type[j] == pc->from_type) && (s->rmass[j] > 0.5 * pc->to_mass)) {
        double delx = xtmp - s->x[3*j], dely = ytmp - s->x[3*j+1], delz = ztmp - s->x[3*j+2];
        double rsq = delx * delx + dely * dely + delz * delz;
        double wfd = kernel_quintic(s->dim, sqrt(rsq) * pc->cutoff);
        double dmass_aux = pc->to_mass * wfd / wtotal;
        dmass[j] += dmass_aux;
        denergy += s->e[j] * dmass_aux;
        for (int d = 0; d < 3; d++) { dmom[d] += s->v[3*j+d] * dmass_aux; dmomest[d] += s->vest[3*j+d] * dmass_aux; }
      }
    }
    (void)denergy;
    int m = s->nlocal - 1;
    s->rmass[m] = pc->to_mass; s->rho[m] = s->rho[i]; s->cv[m] = s->cv[i];
    for (int d = 0; d < 3; d++) { s->v[3*m+d] = dmom[d] / pc->to_mass; s->vest[3*m+d] = dmomest[d] / pc->to_mass; }
    double energy_aux = 0.5 * (s->e[i] - pc->Hwv);
    s->e[i] = energy_aux; s->e[m] = energy_aux;
  }
  fx->pc_nins = nins; fx->pc_nlocal0 = nlocal;
  return 1;
}
/* after comm->reverse_comm_fix(this) (:324): the mass debit and energy renormalisation of the atoms owned before the call (:324-332),
 * then tag_extend / nghost = 0 when ANY rank inserted (:338-350).  itag = first tag for this rank's untagged atoms (Atom::tag_extend, atom.cpp:598-630) */
static void pc_finish(osph_sph *s, ofix *fx, int ninsall, int itag)
{
  double *dmass = s->drho;
  for (int i = 0; i < fx->pc_nlocal0; i++) {
    double mold = s->rmass[i];
    s->rmass[i] -= dmass[i];
    s->e[i] = s->e[i] * mold / s->rmass[i];
    dmass[i] = 0;
  }
  fx->next_reneighbor += fx->pc.nfreq;
  if (ninsall > 0) {
    for (int i = 0; i < s->nlocal; i++) if (s->tag[i] == 0) s->tag[i] = itag++;
    s->nghost = 0;
    for (int k = 0; k < s->nswap; k++) s->sendnum[k] = 0;
    s->ninserted += fx->pc_nins;
  }
}
static int fix_phase_change_pre_exchange(osph_sph *s, ofix *fx)
{
  int did = pc_local(s, fx);
  if (did <= 0) return did;
  comm_reverse_dmass(s); /* :324 (uses the old ghost slots, some now overwritten by new atoms -- as in the reference) */
  int maxtag = 0;
  for (int i = 0; i < s->nlocal; i++) if (s->tag[i] > maxtag) maxtag = s->tag[i];
  pc_finish(s, fx, fx->pc_nins, maxtag + 1);
  return 0;
}

/* ======================================================================
   Verlet
   ====================================================================== */

/* Verlet::force_clear (verlet.cpp:325-371, newton on: nall) + AtomVecMeso*::force_clear (de, drho) */
int osph_force_clear(osph_sph *s)
{
  int nall = s->nlocal + s->nghost;
  memset(s->f, 0, sizeof(double) * 3 * nall); memset(s->de, 0, sizeof(double) * nall); memset(s->drho, 0, sizeof(double) * nall);
  return 0;
}
int osph_pair_compute(osph_sph *s, int slot) { if (slot < 0 || slot >= s->npair) return fail("bad slot"); return pair_compute_slot(s, slot); }
int osph_pair_compute_all(osph_sph *s) { for (int k = 0; k < s->npair; k++) if (pair_compute_slot(s, k)) return -1; return 0; } /* pair_hybrid.cpp:101-109 */
int osph_reverse_comm(osph_sph *s) { comm_reverse(s); return 0; }
int osph_forward_comm(osph_sph *s) { comm_forward(s); return 0; }
int osph_post_force(osph_sph *s)
{
  for (int i = 0; i < s->nfix; i++) {
    if (s->fix[i].kind == FIX_GRAVITY) fix_gravity(s, &s->fix[i]);
    else if (s->fix[i].kind == FIX_SETMESO) fix_setmeso(s, &s->fix[i]);
    else if (s->fix[i].kind == FIX_ENFORCE2D) fix_enforce2d(s, &s->fix[i]);
    else if (s->fix[i].kind == FIX_SETFORCE) fix_setforce(s, &s->fix[i]);
    else if (s->fix[i].kind == FIX_SETMESODE) fix_setmesode(s, &s->fix[i]);
    else if (s->fix[i].kind == FIX_ADDFORCE) fix_addforce(s, &s->fix[i]);
  }
  return 0;
}
int osph_initial_integrate(osph_sph *s)
{ for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == FIX_MESO || s->fix[i].kind == FIX_MESO_STATIONARY) fix_initial_integrate(s, &s->fix[i]); return 0; }
int osph_final_integrate(osph_sph *s)
{ for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == FIX_MESO || s->fix[i].kind == FIX_MESO_STATIONARY) fix_final_integrate(s, &s->fix[i]); return 0; }
int osph_neigh_decide(osph_sph *s, int *rebuild) { *rebuild = neighbor_decide(s); return 0; }

/* Atom::sort, atom.cpp:1555-1650, with the bins of Atom::setup_sort_bins (:1659-1730): owned atoms re-ordered bin by bin,
 * atoms of one bin in their previous order; every per-atom array follows the permutation. */
static int atom_sort(osph_sph *s)
{
  s->nextsort = (s->ntimestep / s->sortfreq) * s->sortfreq + s->sortfreq;
  double binsize = s->userbinsize > 0.0 ? s->userbinsize : 0.5 * s->cutneighmax;
  if (binsize == 0.0) return fail("Atom sorting has bin size = 0.0");
  double bininv = 1.0 / binsize, inv[3]; int nb[3];
  for (int d = 0; d < 3; d++) {
    nb[d] = (int)((s->subhi[d] - s->sublo[d]) * bininv);
    if (s->dim == 2 && d == 2) nb[d] = 1;
    if (nb[d] == 0) nb[d] = 1;
    inv[d] = nb[d] / (s->subhi[d] - s->sublo[d]);
  }
  int nbins = nb[0] * nb[1] * nb[2], n = s->nlocal;
  if (nbins == 1 || n == 0) return 0;
  int *head = malloc(sizeof(int) * nbins), *next = malloc(sizeof(int) * n), *perm = malloc(sizeof(int) * n);
  for (int b = 0; b < nbins; b++) head[b] = -1;
  for (int i = n - 1; i >= 0; i--) {     /* reverse order so that the linked lists run forward (:1586-1603) */
    int c[3];
    for (int d = 0; d < 3; d++) {
      c[d] = (int)((s->x[3*i+d] - s->sublo[d]) * inv[d]);
      if (c[d] < 0) c[d] = 0;
      if (c[d] > nb[d] - 1) c[d] = nb[d] - 1;
    }
    int b = c[2] * nb[1] * nb[0] + c[1] * nb[0] + c[0];
    next[i] = head[b]; head[b] = i;
  }
  int m = 0;
  for (int b = 0; b < nbins; b++) for (int i = head[b]; i >= 0; i = next[i]) perm[m++] = i;     /* new atom m = old atom perm[m] */
#define PERMUTE(T, arr, w) do { T *tmp_ = malloc(sizeof(T) * (w) * n); memcpy(tmp_, arr, sizeof(T) * (w) * n); \
    for (int a_ = 0; a_ < n; a_++) for (int d_ = 0; d_ < (w); d_++) arr[(w)*a_+d_] = tmp_[(w)*perm[a_]+d_]; free(tmp_); } while (0)
  PERMUTE(double, s->x, 3); PERMUTE(double, s->v, 3); PERMUTE(double, s->vest, 3); PERMUTE(double, s->f, 3); PERMUTE(double, s->cg, 3);
  PERMUTE(double, s->rho, 1); PERMUTE(double, s->drho, 1); PERMUTE(double, s->e, 1); PERMUTE(double, s->de, 1);
  PERMUTE(double, s->cv, 1); PERMUTE(double, s->rmass, 1);
  PERMUTE(int, s->type, 1); PERMUTE(int, s->mask, 1); PERMUTE(int, s->tag, 1); PERMUTE(int, s->img, 1);
#undef PERMUTE
  free(head); free(next); free(perm);
  return 0;
}

/* the rebuild branch of Verlet::run, verlet.cpp:240-257 */
int osph_reneighbor(osph_sph *s)
{
  for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == FIX_PHASE_CHANGE) if (fix_phase_change_pre_exchange(s, &s->fix[i])) return -1;
  domain_pbc(s);
  if (s->shrink) {    /* verlet.cpp:244-248: if (domain->box_change) reset_box, comm->setup, neighbor->setup_bins */
    if (domain_reset_box(s)) return -1;
    if (comm_setup(s)) return -1;
    if (setup_bins(s)) return -1;
  }
  /* comm->exchange(): no-op on a 1x1x1 grid (comm_brick.cpp:596 "if (procgrid[dim] == 1) continue") */
  if (s->sortfreq > 0 && s->ntimestep >= s->nextsort) if (atom_sort(s)) return -1;   /* verlet.cpp:251 */
  comm_borders(s);
  return neighbor_build(s);
}

/* Pair::virial_fdotr_compute, pair.cpp:1403-1451 (newton on: owned + ghost atoms, before the reverse halo) */
static void virial_fdotr(osph_sph *s)
{
  int nall = s->nlocal + s->nghost;
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nall; i++) {
    const double *x = &s->x[3*i], *f = &s->f[3*i];
    v[0] += x[0]*f[0]; v[1] += x[1]*f[1]; v[2] += x[2]*f[2]; v[3] += x[0]*f[1]; v[4] += x[0]*f[2]; v[5] += x[1]*f[2];
  }
  memcpy(s->virial, v, sizeof v);
}
int osph_request_virial(osph_sph *s) { s->vir_request = 1; return 0; }
int osph_get_virial(osph_sph *s, double v[6]) { memcpy(v, s->virial, 6 * sizeof(double)); return 0; }

/* Verlet::setup, verlet.cpp:88-142 */
int osph_setup(osph_sph *s)
{
  if (!s->cutneighsq) return fail("setup: b200_neighbor not called");
  if (s->inworld) return fail("a rank of an emulated world is driven by osph_world_setup / osph_world_run");
  domain_pbc(s);
  if (domain_reset_box(s)) return -1;   /* verlet.cpp:102 */
  if (comm_setup(s)) return -1;
  if (setup_bins(s)) return -1;
  if (s->sortfreq > 0) if (atom_sort(s)) return -1;   /* verlet.cpp:106 */
  comm_borders(s);
  if (neighbor_build(s)) return -1;
  s->nbuilds = 0; /* neighbor->ncalls = 0 */
  osph_force_clear(s);
  for (int i = 0; i < s->nfix; i++) /* FixMeso::setup_pre_force, fix_meso.cpp:68-85 */
    if (s->fix[i].kind == FIX_MESO)
      for (int a = 0; a < s->nlocal; a++) if (s->mask[a] & s->fix[i].groupbit) for (int d = 0; d < 3; d++) s->vest[3*a+d] = s->v[3*a+d];
  if (osph_pair_compute_all(s)) return -1;
  if (s->vir_request) { virial_fdotr(s); s->vir_request = 0; }
  comm_reverse(s);
  osph_post_force(s); /* modify->setup: FixGravity::setup -> post_force */
  for (int i = 0; i < s->nfix; i++) if (s->fix[i].kind == FIX_DT_RESET) fix_dt_reset(s, &s->fix[i]);   /* FixDtReset::setup -> end_of_step */
  s->setup_done = 1;
  return 0;
}

/* Verlet::run, verlet.cpp:207-309 */
int osph_run(osph_sph *s, int nsteps)
{
  if (!s->setup_done) return fail("run before setup");
  if (s->inworld) return fail("a rank of an emulated world is driven by osph_world_setup / osph_world_run");
  for (int it = 0; it < nsteps; it++) {
    s->ntimestep++;
    osph_initial_integrate(s);
    if (neighbor_decide(s) == 0) comm_forward(s);
    else if (osph_reneighbor(s)) return -1;
    osph_force_clear(s);
    if (osph_pair_compute_all(s)) return -1;
    if (it == nsteps - 1 && s->vir_request) { virial_fdotr(s); s->vir_request = 0; }
    comm_reverse(s);
    osph_post_force(s);
    osph_final_integrate(s);
    for (int i = 0; i < s->nfix; i++)   /* modify->end_of_step */
      if (s->fix[i].kind == FIX_DT_RESET && s->ntimestep % s->fix[i].nevery == 0) fix_dt_reset(s, &s->fix[i]);
    s->nsteps_done++;
  }
  return 0;
}

/* ======================================================================
   P ranks in one process (test infrastructure for the multi-GPU path): every rank is an osph_sph of its own with a brick
   sub-domain; the collective steps of CommBrick run in lock step over the array of ranks, reading the sender's arrays
   directly where MPI would move a buffer.  Restated for maxneed = 1 (one ghost layer; uniform bricks or the non-uniform cuts of
   `balance ... shift`, every brick at least one ghost cutoff long) and no shrink-wrapped faces; fix phase_change runs with one RanPark
   stream per rank (w_phase_change).
   ====================================================================== */

/* CommBrick::setup, comm_brick.cpp:150-386, for this rank's place in the grid */
static int w_comm_setup(osph_sph *s)
{
  s->nswap = 0;
  for (int d = 0; d < 3; d++) {
    const int pg = s->procgrid[d], loc = s->myloc[d];
    int maxneed = (int)(s->cutghost * pg / s->prd[d]) + 1;                 /* :225-227 */
    if (s->dim == 2 && d == 2) maxneed = 0;
    if (!s->periodic[d] && maxneed > pg - 1) maxneed = pg - 1;             /* :229-231 */
    if (pg > 1 && s->subhi[d] - s->sublo[d] < s->cutghost) maxneed = 2;   /* non-uniform bricks (balance shift): the updown() walk, :260-300 */
    if (maxneed > 1) return fail("oracle world: cutghost >= sub-domain length is not restated");
    int sendneed[2] = {maxneed, maxneed};
    if (!s->periodic[d]) {                                                  /* :233-243 */
      int left = loc - 1; if (left < 0) left = pg - 1;
      int right = loc + 1; if (right == pg) right = 0;
      sendneed[0] = maxneed < pg - left - 1 ? maxneed : pg - left - 1;
      sendneed[1] = maxneed < right ? maxneed : right;
    }
    for (int ineed = 0; ineed < 2 * maxneed; ineed++) {
      int k = s->nswap++;
      s->swapdim[k] = d; s->swappbc[k] = 0;
      if (ineed % 2 == 0) {
        s->sendproc[k] = s->procneigh[d][0]; s->recvproc[k] = s->procneigh[d][1];
        s->slablo[k] = -BIG; s->slabhi[k] = s->sublo[d] + s->cutghost;
        if (loc == 0) s->swappbc[k] = 1;
      } else {
        s->sendproc[k] = s->procneigh[d][1]; s->recvproc[k] = s->procneigh[d][0];
        s->slablo[k] = s->subhi[d] - s->cutghost; s->slabhi[k] = BIG;
        if (loc == pg - 1) s->swappbc[k] = -1;
      }
      s->dosend[k] = (ineed / 2 < sendneed[ineed % 2]);                     /* borders(): sendflag, :741-742 */
    }
  }
  return 0;
}

static void copy_atom(osph_sph *dst, int di, const osph_sph *src, int si)
{
  for (int d = 0; d < 3; d++) {
    dst->x[3*di+d] = src->x[3*si+d]; dst->v[3*di+d] = src->v[3*si+d]; dst->vest[3*di+d] = src->vest[3*si+d];
    dst->f[3*di+d] = src->f[3*si+d]; dst->cg[3*di+d] = src->cg[3*si+d];
  }
  dst->rho[di] = src->rho[si]; dst->drho[di] = src->drho[si]; dst->e[di] = src->e[si]; dst->de[di] = src->de[si];
  dst->cv[di] = src->cv[si]; dst->rmass[di] = src->rmass[si];
  dst->type[di] = src->type[si]; dst->mask[di] = src->mask[si]; dst->tag[di] = src->tag[si]; dst->img[di] = src->img[si];
}

/* CommBrick::exchange, comm_brick.cpp:575-680: atoms that left the sub-domain go to the neighbours of each dimension in turn */
static void w_exchange(osph_sph **R, int n)
{
  const int dimension = R[0]->dim;
  int **out = calloc(n, sizeof(int *)); int *nout = calloc(n, sizeof(int));
  osph_sph *tmp = NULL;        /* the senders' buf_send: atoms leave their rank before anybody receives */
  osph_create(&tmp, -1);
  for (int r = 0; r < n; r++) R[r]->nghost = 0;
  for (int dim = 0; dim < dimension; dim++) {
    int total = 0;
    for (int r = 0; r < n; r++) {
      osph_sph *s = R[r];
      const double lo = s->sublo[dim], hi = s->subhi[dim];
      out[r] = xrealloc(out[r], sizeof(int) * (s->nlocal + 1)); nout[r] = 0;
      int nlocal = s->nlocal, i = 0;
      while (i < nlocal) {
        if (s->x[3*i+dim] < lo || s->x[3*i+dim] >= hi) {
          grow(tmp, total + 1); copy_atom(tmp, total, s, i); out[r][nout[r]++] = total++;
          copy_atom(s, i, s, nlocal - 1);                                   /* avec->copy(nlocal-1,i,1) */
          nlocal--;
        } else i++;
      }
      s->nlocal = nlocal;
    }
    for (int r = 0; r < n; r++) {
      osph_sph *s = R[r];
      if (s->procgrid[dim] == 1) continue;                                  /* the leavers are lost, as in the reference */
      const double lo = s->sublo[dim], hi = s->subhi[dim];
      const int from[2] = {s->procneigh[dim][1], s->procneigh[dim][0]};     /* recv from the right, then (more than 2 procs) from the left */
      for (int side = 0; side < (s->procgrid[dim] > 2 ? 2 : 1); side++) {
        const int q = from[side];
        for (int k = 0; k < nout[q]; k++) {
          const double value = tmp->x[3*out[q][k]+dim];
          if (value >= lo && value < hi) { grow(s, s->nlocal + 1); copy_atom(s, s->nlocal, tmp, out[q][k]); s->nlocal++; }
        }
      }
    }
  }
  for (int r = 0; r < n; r++) free(out[r]);
  free(out); free(nout); osph_destroy(tmp);
}

/* CommBrick::borders, comm_brick.cpp:696-864 */
static void w_borders(osph_sph **R, int n)
{
  int *nlast = calloc(n, sizeof(int));
  for (int r = 0; r < n; r++) R[r]->nghost = 0;
  int iswap = 0; const int nswap = R[0]->nswap;
  while (iswap < nswap) {
    const int dim = R[0]->swapdim[iswap];
    for (int r = 0; r < n; r++) nlast[r] = R[r]->nlocal + R[r]->nghost;
    for (int half = 0; half < 2 && iswap < nswap && R[0]->swapdim[iswap] == dim; half++, iswap++) {
      for (int r = 0; r < n; r++) {                                          /* everybody lists what it sends */
        osph_sph *s = R[r];
        const double lo = s->slablo[iswap], hi = s->slabhi[iswap];
        int nsend = 0;
        if (s->dosend[iswap])
          for (int i = 0; i < nlast[r]; i++)
            if (s->x[3*i+dim] >= lo && s->x[3*i+dim] <= hi) {
              if (nsend == s->maxsend[iswap]) { s->maxsend[iswap] = nsend * 2 + 1024; s->sendlist[iswap] = xrealloc(s->sendlist[iswap], sizeof(int) * s->maxsend[iswap]); }
              s->sendlist[iswap][nsend++] = i;
            }
        s->sendnum[iswap] = nsend;
      }
      for (int r = 0; r < n; r++) {                                          /* ... and unpacks what its recvproc sent */
        osph_sph *s = R[r]; const osph_sph *src = R[s->recvproc[iswap]];
        const int nrecv = src->sendnum[iswap], first = s->nlocal + s->nghost;
        grow(s, first + nrecv);
        src = R[s->recvproc[iswap]];
        const double shift = src->swappbc[iswap] * src->prd[dim];
        const int imgmul = dim == 0 ? 1 : (dim == 1 ? 3 : 9);
        for (int k = 0; k < nrecv; k++) {
          const int j = src->sendlist[iswap][k], g = first + k;
          for (int d = 0; d < 3; d++) {
            s->x[3*g+d] = (d == dim && src->swappbc[iswap]) ? src->x[3*j+d] + shift : src->x[3*j+d];
            s->cg[3*g+d] = src->cg[3*j+d]; s->vest[3*g+d] = src->vest[3*j+d];
            if (s->ghost_velocity) s->v[3*g+d] = src->v[3*j+d];
          }
          s->tag[g] = src->tag[j]; s->type[g] = src->type[j]; s->mask[g] = src->mask[j];
          s->rho[g] = src->rho[j]; s->rmass[g] = src->rmass[j]; s->e[g] = src->e[j]; s->cv[g] = src->cv[j];
          s->img[g] = src->img[j] + src->swappbc[iswap] * imgmul;
        }
        s->firstrecv[iswap] = first; s->nghost += nrecv;
      }
    }
  }
  free(nlast);
}

/* CommBrick::forward_comm, comm_brick.cpp:444-506 */
static void w_forward(osph_sph **R, int n)
{
  for (int iswap = 0; iswap < R[0]->nswap; iswap++) {
    const int dim = R[0]->swapdim[iswap];
    for (int r = 0; r < n; r++) {
      osph_sph *s = R[r]; const osph_sph *src = R[s->recvproc[iswap]];
      const double shift = src->swappbc[iswap] * src->prd[dim];
      for (int k = 0; k < src->sendnum[iswap]; k++) {
        const int j = src->sendlist[iswap][k], g = s->firstrecv[iswap] + k;
        for (int d = 0; d < 3; d++) {
          s->x[3*g+d] = (d == dim && src->swappbc[iswap]) ? src->x[3*j+d] + shift : src->x[3*j+d];
          s->vest[3*g+d] = src->vest[3*j+d];
          if (s->multiphase) s->cg[3*g+d] = src->cg[3*j+d];
          if (s->ghost_velocity) s->v[3*g+d] = src->v[3*j+d];
        }
        s->rho[g] = src->rho[j]; s->e[g] = src->e[j];
        if (s->multiphase) s->rmass[g] = src->rmass[j];
      }
    }
  }
}
/* CommBrick::reverse_comm, comm_brick.cpp:513-560: swaps in reverse order; the ghost of swap k on rank r adds into its sender */
static void w_reverse(osph_sph **R, int n)
{
  for (int iswap = R[0]->nswap - 1; iswap >= 0; iswap--)
    for (int r = 0; r < n; r++) {
      osph_sph *s = R[r]; osph_sph *src = R[s->recvproc[iswap]];
      for (int k = 0; k < src->sendnum[iswap]; k++) {
        const int j = src->sendlist[iswap][k], g = s->firstrecv[iswap] + k;
        for (int d = 0; d < 3; d++) src->f[3*j+d] += s->f[3*g+d];
        src->drho[j] += s->drho[g]; src->de[j] += s->de[g];
      }
    }
}
/* CommBrick::forward_comm_pair with PairSPHRhoSum::pack_forward_comm, comm_brick.cpp:871-904 */
static void w_forward_rho(osph_sph **R, int n)
{
  for (int iswap = 0; iswap < R[0]->nswap; iswap++)
    for (int r = 0; r < n; r++) {
      osph_sph *s = R[r]; const osph_sph *src = R[s->recvproc[iswap]];
      for (int k = 0; k < src->sendnum[iswap]; k++) s->rho[s->firstrecv[iswap] + k] = src->rho[src->sendlist[iswap][k]];
    }
}

static int w_check(osph_sph **R, int n)
{
  if (n < 1) return fail("world: no ranks");
  for (int r = 0; r < n; r++) {
    if (!R[r]->inworld || R[r]->me != r) return fail("world: rank r must have been given osph_comm_init(world, r, ...)");
    if (!R[r]->cutneighsq) return fail("world setup: b200_neighbor not called");
    if (R[r]->shrink) return fail("world: shrink-wrapped boundaries are not restated for P > 1");
  }
  return 0;
}
/* pair_hybrid.cpp:101-109 over all ranks, sub-style by sub-style (PairSPHRhoSum ends with its forward_comm_pair) */
static int w_pair_compute_all(osph_sph **R, int n)
{
  for (int k = 0; k < R[0]->npair; k++) {
    for (int r = 0; r < n; r++) if (pair_compute_slot(R[r], k)) return -1;
    if (R[0]->pair[k].style == B200_PAIR_RHOSUM) w_forward_rho(R, n);
  }
  return 0;
}
/* FixPhaseChange::pre_exchange on every rank (modify->pre_exchange, verlet.cpp:241): every rank walks ITS copy of the RanPark stream (all
 * constructed from the same seed, fix_phase_change.cpp:107) over its own atoms, comm->reverse_comm_fix adds the ghosts' dmass to their
 * owners across the ranks (comm_brick.cpp:930-965 with FixPhaseChange::pack/unpack_reverse_comm :518-537), Atom::tag_extend numbers the
 * new atoms rank after rank */
static int w_phase_change(osph_sph **R, int n)
{
  for (int f = 0; f < R[0]->nfix; f++) {
    if (R[0]->fix[f].kind != FIX_PHASE_CHANGE) continue;
    int due = 0;
    for (int r = 0; r < n; r++) { int d = pc_local(R[r], &R[r]->fix[f]); if (d < 0) return -1; due |= d; }
    if (!due) continue;
    for (int iswap = R[0]->nswap - 1; iswap >= 0; iswap--)
      for (int r = 0; r < n; r++) {
        osph_sph *s = R[r]; osph_sph *src = R[s->recvproc[iswap]];
        for (int k = 0; k < src->sendnum[iswap]; k++) src->drho[src->sendlist[iswap][k]] += s->drho[s->firstrecv[iswap] + k];
      }
    int ninsall = 0, maxtag = 0;
    for (int r = 0; r < n; r++) {
      ninsall += R[r]->fix[f].pc_nins;
      for (int i = 0; i < R[r]->nlocal; i++) if (R[r]->tag[i] > maxtag) maxtag = R[r]->tag[i];
    }
    int itag = maxtag + 1;
    for (int r = 0; r < n; r++) { pc_finish(R[r], &R[r]->fix[f], ninsall, itag); itag += R[r]->fix[f].pc_nins; }
  }
  return 0;
}
static int w_rebuild(osph_sph **R, int n)
{
  if (w_phase_change(R, n)) return -1;
  for (int r = 0; r < n; r++) domain_pbc(R[r]);
  w_exchange(R, n);
  for (int r = 0; r < n; r++)      /* verlet.cpp:251: every rank sorts its own atoms in the bins of its sub-domain (atom.cpp:1555-1730) */
    if (R[r]->sortfreq > 0 && R[r]->ntimestep >= R[r]->nextsort) if (atom_sort(R[r])) return -1;
  w_borders(R, n);
  for (int r = 0; r < n; r++) if (neighbor_build(R[r])) return -1;
  return 0;
}

/* Verlet::setup, verlet.cpp:88-142, on every rank */
int osph_world_setup(osph_sph **R, int n)
{
  if (w_check(R, n)) return -1;
  for (int r = 0; r < n; r++) { domain_pbc(R[r]); if (w_comm_setup(R[r])) return -1; if (setup_bins(R[r])) return -1; }
  w_exchange(R, n);
  for (int r = 0; r < n; r++) if (R[r]->sortfreq > 0) if (atom_sort(R[r])) return -1;      /* verlet.cpp:106 */
  w_borders(R, n);
  for (int r = 0; r < n; r++) {
    osph_sph *s = R[r];
    if (neighbor_build(s)) return -1;
    s->nbuilds = 0;
    osph_force_clear(s);
    for (int i = 0; i < s->nfix; i++)
      if (s->fix[i].kind == FIX_MESO)
        for (int a = 0; a < s->nlocal; a++) if (s->mask[a] & s->fix[i].groupbit) for (int d = 0; d < 3; d++) s->vest[3*a+d] = s->v[3*a+d];
  }
  if (w_pair_compute_all(R, n)) return -1;
  w_reverse(R, n);
  for (int r = 0; r < n; r++) { osph_post_force(R[r]); R[r]->setup_done = 1; }
  return 0;
}

/* Verlet::run, verlet.cpp:207-309, on every rank */
int osph_world_run(osph_sph **R, int n, int nsteps)
{
  if (w_check(R, n)) return -1;
  for (int r = 0; r < n; r++) if (!R[r]->setup_done) return fail("run before setup");
  for (int it = 0; it < nsteps; it++) {
    int rebuild = 0;
    for (int r = 0; r < n; r++) { R[r]->ntimestep++; osph_initial_integrate(R[r]); }
    for (int r = 0; r < n; r++) rebuild |= neighbor_decide(R[r]);            /* MPI_Allreduce MAX, neighbor.cpp:1407 */
    if (!rebuild) w_forward(R, n);
    else if (w_rebuild(R, n)) return -1;
    for (int r = 0; r < n; r++) osph_force_clear(R[r]);
    if (w_pair_compute_all(R, n)) return -1;
    w_reverse(R, n);
    for (int r = 0; r < n; r++) { osph_post_force(R[r]); osph_final_integrate(R[r]); R[r]->nsteps_done++; }
  }
  return 0;
}

/* ---------------------------------------------------------------------- */

static int cmp_ent(const void *a, const void *b)
{ const int *p = a, *q = b; if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1; return p[1] - q[1]; }

int osph_get_neighbor_list(osph_sph *s, int nlocal, int *numneigh, long long nentries, int *jtag, int *jimage)
{
  if (nlocal != s->nlocal) return fail("get_neighbor_list: nlocal mismatch");
  long long tot = 0;
  for (int i = 0; i < nlocal; i++) { numneigh[i] = s->numneigh[i]; tot += s->numneigh[i]; }
  if (!jtag) return 0;
  if (nentries < tot) return fail("get_neighbor_list: buffer too small");
  long long o = 0;
  for (int i = 0; i < nlocal; i++) {
    int n = s->numneigh[i]; int *tmp = malloc(sizeof(int) * 2 * (n + 1));
    for (int k = 0; k < n; k++) { int j = s->neigh[s->firstneigh[i] + k]; tmp[2*k] = s->tag[j]; tmp[2*k+1] = s->img[j]; }
    qsort(tmp, n, 2 * sizeof(int), cmp_ent);
    for (int k = 0; k < n; k++) { jtag[o + k] = tmp[2*k]; jimage[o + k] = tmp[2*k+1]; }
    free(tmp); o += n;
  }
  return 0;
}

int osph_get_counters(osph_sph *s, long long c[8])
{ c[0] = 0; c[1] = s->nbuilds; c[2] = s->nsteps_done; c[3] = s->maxneighseen; c[4] = s->nghost; c[5] = 0; c[6] = s->ninserted; c[7] = s->ndanger; return 0; }
int osph_set_timing(osph_sph *s, int on) { (void)s; (void)on; return 0; }
int osph_get_timers(osph_sph *s, int n, double *ms, long long *calls) { (void)s; for (int i = 0; i < n; i++) { ms[i] = 0; calls[i] = 0; } return 0; }
const char *osph_timer_name(int i) { (void)i; return ""; }
int osph_sync(osph_sph *s) { (void)s; return 0; }
