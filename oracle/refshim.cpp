// refshim.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin extern "C" window into the *unmodified* reference LAMMPS objects
// (oracle/_ref/liblammps_ref.so, compiled from /root/reference by oracle/Makefile)
// for the things its C library API (src/library.cpp) does not expose: ghost
// counts, colorgradient, tags, and the built full neighbor list.  It contains
// no SPH arithmetic of its own.
#include <mpi.h>
#include "lammps.h"
#include "atom.h"
#define protected public   // Neighbor::cutneighsq is protected; read-only peek
#include "neighbor.h"
#undef protected
#include "neigh_list.h"
#include "neigh_request.h"
#include "input.h"
#include "update.h"
#include "domain.h"
#include "force.h"
#include "pair.h"
#include "comm.h"
#include "timer.h"
#include <cstring>
#include <cstdlib>
using namespace LAMMPS_NS;

extern "C" {

void *refshim_open(int quiet)
{
  int argc = 0;
  static char *argv[8];
  argv[argc++] = (char *)"lammps";
  if (quiet) { argv[argc++] = (char *)"-screen"; argv[argc++] = (char *)"none"; argv[argc++] = (char *)"-log"; argv[argc++] = (char *)"none"; }
  int flag; MPI_Initialized(&flag);
  if (!flag) { int a = 0; char **b = NULL; MPI_Init(&a, &b); }
  LAMMPS *l = new LAMMPS(argc, argv, MPI_COMM_WORLD);
  // The fork never initialises Atom::colorgradient / colorgradient_flag in Atom::Atom
  // (src/atom.cpp:94-95,149), so AtomVecMesoMultiPhase::grow reallocs a garbage pointer
  // unless the heap happens to be zero (fresh lmp_serial process).  Inside a long-lived
  // host process (python) it is not; give the members the value a fresh process sees.
  l->atom->colorgradient = NULL;
  l->atom->colorgradient_flag = 0;
  return (void *)l;
}
void refshim_close(void *p) { delete (LAMMPS *)p; }
void refshim_command(void *p, const char *cmd) { char *s = strdup(cmd); ((LAMMPS *)p)->input->one(s); free(s); }
void refshim_file(void *p, const char *fn) { ((LAMMPS *)p)->input->file(fn); }

int refshim_nlocal(void *p) { return ((LAMMPS *)p)->atom->nlocal; }
int refshim_nghost(void *p) { return ((LAMMPS *)p)->atom->nghost; }
int refshim_ntypes(void *p) { return ((LAMMPS *)p)->atom->ntypes; }
long long refshim_ntimestep(void *p) { return ((LAMMPS *)p)->update->ntimestep; }
double refshim_dt(void *p) { return ((LAMMPS *)p)->update->dt; }
long long refshim_nbuilds(void *p) { return ((LAMMPS *)p)->neighbor->ncalls; }
long long refshim_ndanger(void *p) { return ((LAMMPS *)p)->neighbor->ndanger; }
double refshim_timer(void *p, int which) { return ((LAMMPS *)p)->timer->array[which]; }

void refshim_box(void *p, double *boxlo, double *boxhi, int *periodicity, int *dim)
{
  Domain *d = ((LAMMPS *)p)->domain;
  for (int i = 0; i < 3; i++) { boxlo[i] = d->boxlo[i]; boxhi[i] = d->boxhi[i]; periodicity[i] = d->periodicity[i]; }
  *dim = d->dimension;
}

// per-atom double fields, n = nlocal (+ nghost if with_ghost); returns columns or -1
int refshim_getd(void *p, const char *name, int with_ghost, double *out)
{
  Atom *a = ((LAMMPS *)p)->atom;
  int n = a->nlocal + (with_ghost ? a->nghost : 0);
  double **v3 = NULL; double *v1 = NULL;
  if (!strcmp(name, "x")) v3 = a->x; else if (!strcmp(name, "v")) v3 = a->v; else if (!strcmp(name, "f")) v3 = a->f;
  else if (!strcmp(name, "vest")) v3 = a->vest; else if (!strcmp(name, "colorgradient")) v3 = a->colorgradient;
  else if (!strcmp(name, "rho")) v1 = a->rho; else if (!strcmp(name, "drho")) v1 = a->drho; else if (!strcmp(name, "e")) v1 = a->e;
  else if (!strcmp(name, "de")) v1 = a->de; else if (!strcmp(name, "cv")) v1 = a->cv; else if (!strcmp(name, "rmass")) v1 = a->rmass;
  else if (!strcmp(name, "mass")) { for (int i = 0; i <= a->ntypes; i++) out[i] = a->mass ? a->mass[i] : 0.0; return 1; }
  if (v3) { if (n) memcpy(out, &v3[0][0], sizeof(double) * 3 * n); return 3; }
  if (v1) { memcpy(out, v1, sizeof(double) * n); return 1; }
  return -1;
}
int refshim_geti(void *p, const char *name, int with_ghost, int *out)
{
  Atom *a = ((LAMMPS *)p)->atom;
  int n = a->nlocal + (with_ghost ? a->nghost : 0);
  if (!strcmp(name, "type")) memcpy(out, a->type, sizeof(int) * n);
  else if (!strcmp(name, "mask")) memcpy(out, a->mask, sizeof(int) * n);
  else if (!strcmp(name, "tag")) for (int i = 0; i < n; i++) out[i] = (int)a->tag[i];
  else return -1;
  return 1;
}

// Neighbor tables as LAMMPS computed them (neighbor.cpp:259-282)
void refshim_cutneigh(void *p, double *cutneighsq, double *cutneighmax, double *skin, int *every, int *delay, int *check, double *cutghost)
{
  LAMMPS *l = (LAMMPS *)p; Neighbor *nb = l->neighbor; int n = l->atom->ntypes;
  for (int i = 0; i <= n; i++) for (int j = 0; j <= n; j++) cutneighsq[i * (n + 1) + j] = (i && j) ? nb->cutneighsq[i][j] : 0.0;
  *cutneighmax = nb->cutneighmax; *skin = nb->skin; *every = nb->every; *delay = nb->delay; *check = nb->dist_check;
  *cutghost = l->comm->cutghost[0];
}

// the built (non-skip, non-copy) full list: Neighbor::full_bin output
static NeighList *full_list(LAMMPS *l)
{
  Neighbor *nb = l->neighbor;
  for (int i = 0; i < nb->nlist; i++) {
    if (i >= nb->old_nrequest) break;
    NeighRequest *r = nb->old_requests[i];   // Neighbor::init moves requests -> old_requests
    if (nb->lists[i] && r->full && !r->skip && !r->copy && !r->occasional && nb->lists[i]->buildflag) return nb->lists[i];
  }
  return NULL;
}
// fallback when no sub-style asked for a full list: the built half list
// (half_bin_newton, neigh_half_bin.cpp), to be symmetrised by the caller
static NeighList *half_list(LAMMPS *l)
{
  Neighbor *nb = l->neighbor;
  for (int i = 0; i < nb->nlist; i++) {
    if (i >= nb->old_nrequest) break;
    NeighRequest *r = nb->old_requests[i];
    if (nb->lists[i] && r->half && !r->skip && !r->copy && !r->occasional && !r->half_from_full && nb->lists[i]->buildflag) return nb->lists[i];
  }
  return NULL;
}
// force->pair->virial of the last step on which the virial was tallied (thermo steps), xx yy zz xy xz yz
void refshim_virial(void *p, double *v) { memcpy(v, ((LAMMPS *)p)->force->pair->virial, 6 * sizeof(double)); }
int refshim_has_full(void *p) { return full_list((LAMMPS *)p) != NULL; }
// numneigh[nlocal]; if j != NULL also fills entries (local indices incl. ghosts) in list order; returns total or -1
long long refshim_neigh_full(void *p, int *numneigh, long long nentries, int *j)
{
  LAMMPS *l = (LAMMPS *)p; NeighList *list = full_list(l);
  if (!list) list = half_list(l);
  if (!list) return -1;
  long long tot = 0;
  for (int ii = 0; ii < list->inum; ii++) { int i = list->ilist[ii]; numneigh[i] = list->numneigh[i]; tot += list->numneigh[i]; }
  if (!j) return tot;
  if (nentries < tot) return -2;
  long long o = 0;
  for (int i = 0; i < l->atom->nlocal; i++) {   // ilist[ii] == ii for full_bin
    for (int k = 0; k < list->numneigh[i]; k++) j[o + k] = list->firstneigh[i][k] & NEIGHMASK;
    o += list->numneigh[i];
  }
  return tot;
}
}
