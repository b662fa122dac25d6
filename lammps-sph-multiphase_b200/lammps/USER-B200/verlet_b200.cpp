/* ----------------------------------------------------------------------
   run_style verlet/b200 -- see verlet_b200.h.  One engine instance (one GPU) per MPI rank.
------------------------------------------------------------------------- */
#include "string.h"
#include "math.h"
#include "verlet_b200.h"
#include "pair_sph_b200.h"
#include "fix_b200.h"
#include "pair_hybrid.h"
#include "neighbor.h"
#include "neigh_request.h"
#include "domain.h"
#include "comm.h"
#include "atom.h"
#include "atom_vec.h"
#include "force.h"
#include "pair.h"
#include "modify.h"
#include "fix.h"
#include "output.h"
#include "update.h"
#include "timer.h"
#include "memory.h"
#include "error.h"

using namespace LAMMPS_NS;

VerletB200::VerletB200(LAMMPS *lmp, int narg, char **arg) : Verlet(lmp, narg, arg), h(NULL), dtfix(NULL) {}

VerletB200::~VerletB200() { if (h) b200_destroy(h); }

// Domain::small (domain.h:148) is protected; the engine needs it to re-fit shrink-wrapped faces exactly as Domain::reset_box does
namespace {
struct DomainPeek : public Domain {
  static const double *small_of(const Domain *d) { return d->*(&DomainPeek::small); }
};
// Atom::userbinsize (atom.h, `atom_modify sort Nfreq binsize`) is protected as well
struct AtomPeek : public Atom {
  static double userbinsize_of(const Atom *a) { return a->*(&AtomPeek::userbinsize); }
};
// Comm::mode / bordergroup (comm.h:102-103, protected): comm_modify mode multi | group ID
struct CommPeek : public Comm {
  static int mode_of(const Comm *c) { return c->*(&CommPeek::mode); }
  static int bordergroup_of(const Comm *c) { return c->*(&CommPeek::bordergroup); }
};
// PairHybrid::nmap / map (pair_hybrid.h:59-60, protected): which sub-styles a type pair is assigned to
struct HybridPeek : public PairHybrid {
  // !ijskip of sub-style m's neighbor request, exactly as PairHybrid::init_style forms it (pair_hybrid.cpp:452-466): the pair is assigned
  // to m, OR it is assigned to nothing while I,I and J,J both belong to m alone.  The second clause is the reference's "mixing will assign
  // this pair" case, and it also catches `pair_coeff I J none` between two types of the same sub-style: such a pair is NOT skipped -- it is
  // kept out of the list only by its zero neighbor cutoff (PairHybrid::init_one returns 0), so two atoms of these types at the SAME
  // position (rsq = 0 <= 0) do interact, with the sub-style's own coefficients (cavity_flow.lmp creates such twins where walls and driver overlap)
  static bool maps(const PairHybrid *p, int i, int j, int m)
  {
    int **nmap = p->*(&HybridPeek::nmap); int ***map = p->*(&HybridPeek::map);
    if (i > j) { int s = i; i = j; j = s; }
    for (int k = 0; k < nmap[i][j]; k++) if (map[i][j][k] == m) return true;
    return nmap[i][j] == 0 && nmap[i][i] == 1 && map[i][i][0] == m && nmap[j][j] == 1 && map[j][j][0] == m;
  }
};
}

void VerletB200::check(int rc) { if (rc < 0) error->all(FLERR, b200_last_error()); }

void VerletB200::init()
{
  Verlet::init();
  // several MPI ranks = one engine instance (one GPU) per rank, LAMMPS' own brick decomposition handed to b200_comm_init (configure).
  // The engine's halo is CommBrick with one ghost layer per swap: a multi-layer setup (comm_brick.cpp:150-170, maxneed > 1) and
  // the tiled layout (comm_style tiled, comm_tiled.cpp) are refused here rather than approximated.
  if (comm->nprocs > 1) {
    if (comm->style != 0) error->all(FLERR, "run_style verlet/b200 supports comm_style brick");
    if (comm->layout > 1) error->all(FLERR, "run_style verlet/b200 supports brick layouts (uniform, or non-uniform from balance shift)");      // enum{LAYOUT_UNIFORM,LAYOUT_NONUNIFORM,LAYOUT_TILED}, comm.cpp:44
    for (int d = 0; d < domain->dimension; d++)
      if (comm->cutghost[d] > domain->subhi[d] - domain->sublo[d])
        error->all(FLERR, "run_style verlet/b200: the ghost cutoff exceeds the sub-domain (more than one ghost layer per swap)");
  }
  // per-atom energy / virial tallies (Pair::ev_tally into eatom / vatom, pair.cpp:867-1000) are not formed by the engine, only the
  // global pair virial is: a compute that reads them (stress/atom, pe/atom, ...; Integrate::ev_setup collected them) is refused
  if (nelist_atom || nvlist_atom)
    error->all(FLERR, "run_style verlet/b200: per-atom energy / virial tallies (compute pe/atom, stress/atom, ...) are not computed by the engine");
  if (CommPeek::mode_of(comm) || CommPeek::bordergroup_of(comm))
    error->all(FLERR, "run_style verlet/b200 supports comm_modify mode single without a border group (one ghost cutoff for all atoms)");
  if (!force->newton_pair) error->all(FLERR, "run_style verlet/b200 requires newton on");
  if (domain->triclinic) error->all(FLERR, "run_style verlet/b200 supports orthogonal boxes");
  if (!atom->rho_flag || !atom->e_flag) error->all(FLERR, "run_style verlet/b200 requires atom_style meso or meso/multiphase");
}

/* push everything LAMMPS parsed across the C-ABI (tables only, verbatim from the host objects) */
void VerletB200::configure()
{
  if (!h) {
    // one rank per GPU: ranks of a node take the devices round robin (the launcher's binding, as -pk gpu does in lib/gpu)
    int ndev = 1;
    if (const char *e = getenv("B200_DEVICES_PER_NODE")) ndev = atoi(e) > 0 ? atoi(e) : 1;
    check(b200_create(&h, comm->nprocs > 1 ? comm->me % ndev : 0));
    if (comm->nprocs > 1) {
      // NCCL rendezvous over MPI: rank 0 draws the unique id, everybody receives it (ncclGetUniqueId / ncclCommInitRank)
      char id[128];
      memset(id, 0, sizeof id);
      if (comm->me == 0) check(b200_comm_unique_id(id));
      MPI_Bcast(id, 128, MPI_CHAR, 0, world);
      int procneigh[6];
      for (int d = 0; d < 3; d++) { procneigh[2 * d] = comm->procneigh[d][0]; procneigh[2 * d + 1] = comm->procneigh[d][1]; }
      check(b200_comm_init(h, comm->nprocs, comm->me, comm->procgrid, comm->myloc, procneigh, id));
    }
  }
  // the engine restates the binned builds (full_bin / half_bin_newton) over all atoms and type pairs (checked here, after
  // Neighbor::init has digested the neigh_modify settings: LAMMPS::init runs it after Integrate::init)
  if (neighbor->style != 1) error->all(FLERR, "run_style verlet/b200 supports neighbor style bin");       // enum{NSQ,BIN,MULTI}, neighbor.cpp:50
  if (neighbor->exclude_setting()) error->all(FLERR, "run_style verlet/b200: neigh_modify exclude is not supported (use pair_coeff I J none)");
  if (neighbor->includegroup) error->all(FLERR, "run_style verlet/b200: neigh_modify include is not supported");
  // the host holds no ghosts and builds no neighbor lists under verlet/b200 (the /b200 pair styles request none): a compute, fix or
  // command that asked Neighbor for one (compute rdf, coord/atom, ...; Neighbor::init moved the requests to old_requests) would read garbage
  for (int i = 0; i < neighbor->old_nrequest; i++) {
    NeighRequest *rq = neighbor->old_requests[i];
    if (rq->compute || rq->fix || rq->command)
      error->all(FLERR, "run_style verlet/b200: a compute / fix / command that needs a host neighbor list is not supported");
  }
  int n = atom->ntypes;
  int multiphase = atom->rmass_flag ? 1 : 0;
  check(b200_domain(h, domain->dimension, domain->boxlo, domain->boxhi, domain->periodicity, domain->sublo, domain->subhi));
  if (domain->nonperiodic == 2) {                 // boundary s / m: the engine owns the box between output steps
    const double minbox[6] = {domain->minxlo, domain->minxhi, domain->minylo, domain->minyhi, domain->minzlo, domain->minzhi};
    check(b200_boundary(h, &domain->boundary[0][0], DomainPeek::small_of(domain), minbox));
  }
  check(b200_atom_style(h, multiphase, n, atom->mass));
  // Neighbor::init, neighbor.cpp:259-282 (cutneighsq itself is protected there)
  std::vector<double> cn((n + 1) * (n + 1), 0.0);
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      double cutoff = sqrt(force->pair->cutsq[i][j]);
      double cut = cutoff + (cutoff > 0.0 ? neighbor->skin : 0.0);
      cn[i * (n + 1) + j] = cut * cut;
    }
  check(b200_neighbor(h, neighbor->skin, neighbor->every, neighbor->delay, neighbor->dist_check, cn.data(), neighbor->cutneighmax,
                      comm->cutghost[0]));
  check(b200_timestep(h, update->dt, force->ftm2v, update->ntimestep));
  dtfix = NULL;
  for (int i = 0; i < modify->nfix; i++) if (FixDtResetB200 *f = dynamic_cast<FixDtResetB200 *>(modify->fix[i])) dtfix = f;
  check(b200_set_time(h, update->atime, update->atimestep, dtfix ? dtfix->laststep : update->ntimestep));
  check(b200_comm_modify(h, comm->ghost_velocity));
  check(b200_atom_modify(h, atom->sortfreq, AtomPeek::userbinsize_of(atom)));      // the engine re-numbers its local indices when Atom::sort would (verlet.cpp:251)

  // pair sub-styles in PairHybrid::compute order (pair_hybrid.cpp:101-109)
  check(b200_pair_clear(h));
  int nsub = 1; Pair **subs = &force->pair;
  PairHybrid *hyb = dynamic_cast<PairHybrid *>(force->pair);
  if (hyb) { nsub = hyb->nstyles; subs = hyb->styles; }
  for (int m = 0; m < nsub; m++) {
    B200PairShell *shell = dynamic_cast<B200PairShell *>(subs[m]);
    if (!shell) error->all(FLERR, "run_style verlet/b200: every pair (sub-)style must be a /b200 style");
    b200_pair_desc d; std::vector<std::vector<double> > ds; std::vector<std::vector<int> > is;
    ds.reserve(16); is.reserve(4);
    shell->b200_describe(d, ds, is);
    // under pair hybrid the type pairs a sub-style computes are those its neighbor request does not skip (HybridPeek::maps), not the
    // sub-style's own setflag: `pair_coeff * * A ...` followed by `pair_coeff 1 2 B ...` leaves A's setflag set for (1,2)
    std::vector<int> mapped;
    if (hyb) {
      mapped.assign((n + 1) * (n + 1), 0);
      for (int i = 1; i <= n; i++)
        for (int j = 1; j <= n; j++) mapped[i * (n + 1) + j] = HybridPeek::maps(hyb, i, j, m) ? 1 : 0;
      d.mapped = mapped.data();
    }
    check(b200_pair_add(h, &d));
  }

  // fixes in Modify order.  A fix with per-step hooks but no /b200 variant cannot run device-resident: it is refused, with one
  // exception -- read-only END_OF_STEP fixes (fix print, fix ave/*): the run is cut into segments that end on their `nevery`,
  // the host arrays are refreshed there and modify->end_of_step() is called exactly as Verlet::run does (verlet.cpp:300).
  // b200_fix_clear keeps the state of an unchanged fix phase_change (next step, RNG position) across `run` commands, as
  // FixPhaseChange itself does (fix_phase_change.cpp:116,345).
  check(b200_fix_clear(h));
  host_every.clear();
  for (int i = 0; i < modify->nfix; i++) {
    Fix *f = modify->fix[i];
    B200FixShell *shell = dynamic_cast<B200FixShell *>(f);
    if (shell) { check(shell->b200_register(h)); continue; }
    int mask = modify->fmask[i];
    const int stepping = FixConst::INITIAL_INTEGRATE | FixConst::POST_INTEGRATE | FixConst::PRE_EXCHANGE | FixConst::PRE_NEIGHBOR |
                         FixConst::PRE_FORCE | FixConst::POST_FORCE | FixConst::FINAL_INTEGRATE |
                         FixConst::INITIAL_INTEGRATE_RESPA | FixConst::POST_INTEGRATE_RESPA | FixConst::PRE_FORCE_RESPA |
                         FixConst::POST_FORCE_RESPA | FixConst::FINAL_INTEGRATE_RESPA;
    bool readonly_eos = strcmp(f->style, "print") == 0 || strncmp(f->style, "ave/", 4) == 0;
    if ((mask & stepping) || ((mask & FixConst::END_OF_STEP) && !readonly_eos)) {
      char msg[256];
      sprintf(msg, "run_style verlet/b200: fix %s (%s) has no /b200 variant", f->id, f->style);
      error->all(FLERR, msg);
    }
    if (mask & FixConst::END_OF_STEP) host_every.push_back(f->nevery > 0 ? f->nevery : 1);
  }
}

void VerletB200::upload()
{
  b200_atoms a; memset(&a, 0, sizeof a);
  int nl = atom->nlocal;
  a.x = nl ? &atom->x[0][0] : NULL; a.v = nl ? &atom->v[0][0] : NULL; a.vest = nl ? &atom->vest[0][0] : NULL; a.f = nl ? &atom->f[0][0] : NULL;
  a.rho = atom->rho; a.drho = atom->drho; a.e = atom->e; a.de = atom->de; a.cv = atom->cv; a.rmass = atom->rmass;
  a.colorgradient = (atom->rmass_flag && nl) ? &atom->colorgradient[0][0] : NULL;
  a.type = atom->type; a.mask = atom->mask;
  std::vector<int> tag(nl);
  for (int i = 0; i < nl; i++) tag[i] = (int)atom->tag[i];
  a.tag = tag.data();
  check(b200_set_atoms(h, nl, &a));
}

void VerletB200::download()
{
  int nl, ng;
  check(b200_get_natoms(h, &nl, &ng));
  if (nl != atom->nlocal) {                    // fix phase_change/b200 created atoms, or atoms migrated between the ranks' engines
    while (nl > atom->nmax) atom->avec->grow(0);
    // what the engine does not carry starts from create_atom's defaults for atoms this rank has not held before (atom_vec_meso.cpp:760-790);
    // image flags of wrapped atoms are not tracked by the engine (limits)
    for (int i = atom->nlocal; i < nl; i++)
      atom->image[i] = ((imageint) IMGMAX << IMG2BITS) | ((imageint) IMGMAX << IMGBITS) | IMGMAX;
    atom->nlocal = nl; atom->nghost = 0;
    bigint nblocal = nl;
    MPI_Allreduce(&nblocal, &atom->natoms, 1, MPI_LMP_BIGINT, MPI_SUM, world);
  }
  b200_atoms a; memset(&a, 0, sizeof a);
  a.x = &atom->x[0][0]; a.v = &atom->v[0][0]; a.vest = &atom->vest[0][0]; a.f = &atom->f[0][0];
  a.rho = atom->rho; a.drho = atom->drho; a.e = atom->e; a.de = atom->de; a.cv = atom->cv; a.rmass = atom->rmass;
  a.colorgradient = atom->rmass_flag ? &atom->colorgradient[0][0] : NULL;
  a.type = atom->type; a.mask = atom->mask;
  std::vector<int> tag(nl);
  a.tag = tag.data();
  check(b200_get_atoms(h, nl, &a));
  for (int i = 0; i < nl; i++) atom->tag[i] = tag[i];
  atom->nghost = 0;
  // An atom beyond a fixed (f) face: the reference drops it at its next reneighboring (CommBrick::exchange keeps what lies inside the
  // sub-box, comm_brick.cpp:596-650) and thermo then stops with "Lost atoms" (thermo.cpp lost_check); the engine keeps such an atom in
  // its boundary cell.  The two would part ways silently, so the run stops here with a message instead.
  {
    bigint nout = 0;
    for (int d = 0; d < domain->dimension; d++) {
      const bool flo = domain->boundary[d][0] == 1, fhi = domain->boundary[d][1] == 1;      // 0 p, 1 f, 2 s, 3 m (domain.h)
      if (!flo && !fhi) continue;
      for (int i = 0; i < nl; i++)
        if ((flo && atom->x[i][d] < domain->boxlo[d]) || (fhi && atom->x[i][d] >= domain->boxhi[d])) nout++;
    }
    bigint nall;
    MPI_Allreduce(&nout, &nall, 1, MPI_LMP_BIGINT, MPI_SUM, world);
    if (nall) {
      char msg[256];
      sprintf(msg, "run_style verlet/b200: " BIGINT_FORMAT " atom coordinate(s) beyond a fixed box face at step " BIGINT_FORMAT
              " (the reference would lose these atoms at its next reneighboring: Lost atoms)", nall, update->ntimestep);
      error->all(FLERR, msg);
    }
  }
  if (atom->map_style) { atom->map_init(); atom->map_set(); }
  long long c[8];
  b200_get_counters(h, c);
  neighbor->ncalls = c[1]; neighbor->ndanger = c[7];
  if (domain->nonperiodic == 2) {                 // thermo volume, dumps: the box as the last rebuild left it
    check(b200_get_box(h, domain->boxlo, domain->boxhi));
    domain->set_global_box();
    domain->set_local_box();
  }
}

/* update->dt and, under fix dt/reset/b200, the elapsed-time bookkeeping of Update::update_time (update.cpp:480-484; thermo keyword
   `time`, thermo.cpp:1500) and FixDtReset::laststep -- the engine advanced them on the device exactly as end_of_step does on the host */
void VerletB200::pull_time()
{
  double dtnow;
  check(b200_get_timestep(h, &dtnow));
  update->dt = dtnow;
  if (dtfix) {
    double atime; long long atimestep, laststep;
    check(b200_get_time(h, &atime, &atimestep, &laststep));
    update->atime = atime; update->atimestep = atimestep; dtfix->laststep = laststep;
  }
}

/* Verlet::setup, verlet.cpp:88-142 */
void VerletB200::setup()
{
  if (comm->me == 0 && screen) fprintf(screen, "Setting up run (B200 engine: %s) ...\n", b200_version());
  update->setupflag = 1;
  atom->setup();
  modify->setup_pre_exchange();
  domain->pbc();
  domain->reset_box();
  comm->setup();
  comm->exchange();
  if (atom->sortfreq > 0) atom->sort();
  atom->nghost = 0;
  configure();
  upload();
  ev_set(update->ntimestep);
  if (vflag) check(b200_request_virial(h));       // thermo output of step 0 needs the pair virial (Pair::virial_fdotr_compute)
  check(b200_setup(h));
  pull_time();                                    // FixDtReset::setup may already have changed the timestep
  download();
  if (vflag && force->pair) check(b200_get_virial(h, force->pair->virial));
  modify->setup(vflag);
  output->setup();
  update->setupflag = 0;
}

void VerletB200::setup_minimal(int flag)
{
  if (flag) setup();
}

/* Verlet::run, verlet.cpp:207-309: device-resident segments between output steps and the steps on which a host-side
   END_OF_STEP fix (fix print, fix ave/...) is due */
void VerletB200::run(int n)
{
  bigint nend = update->ntimestep + n;
  while (update->ntimestep < nend) {
    bigint next = output->next < nend ? output->next : nend;
    for (size_t q = 0; q < host_every.size(); q++) {
      bigint due = (update->ntimestep / host_every[q] + 1) * host_every[q];
      if (due < next) next = due;
    }
    if (next <= update->ntimestep) next = update->ntimestep + 1;
    int k = (int)(next - update->ntimestep);
    timer->stamp();
    ev_set(next);                                   // does the step that ends this segment tally the virial? (integrate.cpp:120-150)
    if (vflag) check(b200_request_virial(h));
    check(b200_run(h, k));
    check(b200_sync(h));
    update->ntimestep += k;
    pull_time();                                    // fix dt/reset/b200 changes the timestep on the device
    timer->stamp(TIME_PAIR);
    bool host_due = false;
    for (size_t q = 0; q < host_every.size(); q++) if (update->ntimestep % host_every[q] == 0) host_due = true;
    if (update->ntimestep == output->next || update->ntimestep == nend || host_due) {
      download();
      if (vflag && force->pair) check(b200_get_virial(h, force->pair->virial));
      timer->stamp(TIME_COMM);
    }
    if (host_due) modify->end_of_step();            // the /b200 shells' own end_of_step hooks are no-ops under verlet/b200
    if (update->ntimestep == output->next) {
      ev_set(update->ntimestep);
      output->write(update->ntimestep);
      timer->stamp(TIME_OUTPUT);
    }
  }
}
