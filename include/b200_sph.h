/* b200_sph.h -- C-ABI of the B200-native SPH hot path (libb200sph.so).
 *
 * This is the drop-in boundary for the reference's USER-SPH path.  A LAMMPS
 * `/b200` Pair/Fix/Integrate shell binds exactly these
 * entry points; nothing but plain pointers, sizes and ints crosses it.
 * The only implementation is the sm_100a CUDA library: there is no CPU
 * fallback.  All `file:line` citations are relative to /root/reference/.
 *
 * Conventions (src/GPU prior art, gpu_extra.h:26-64):
 *   - every call returns 0 on success, <0 on error; b200_last_error() gives text
 *   - per-type tables are 1-based, length ntypes+1        (src/pair.h cutsq etc.)
 *   - per-type-pair tables are row-major (ntypes+1)x(ntypes+1), 1-based
 *   - per-atom vectors are LAMMPS' contiguous AoS blocks: x = &atom->x[0][0]
 *     is [n][3] doubles (src/memory.h:124-137); the library never keeps host
 *     pointers after a call returns
 *   - all physics is fp64, indices int32
 */
#ifndef B200_SPH_H
#define B200_SPH_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_sph b200_sph; /* one engine instance per MPI rank / GPU */

/* ---- pair sub-styles (src/USER-SPH/pair_sph_*.h PairStyle() names) -------- */
enum {
  B200_PAIR_RHOSUM = 1,             /* sph/rhosum                    pair_sph_rhosum.cpp:66-204 */
  B200_PAIR_RHOSUM_MULTIPHASE = 2,  /* sph/rhosum/multiphase         pair_sph_rhosum_multiphase.cpp:68-174 */
  B200_PAIR_TAITWATER = 3,          /* sph/taitwater                 pair_sph_taitwater.cpp:53-200 */
  B200_PAIR_TAITWATER_MORRIS = 4,   /* sph/taitwater/morris          pair_sph_taitwater_morris.cpp:52-200 */
  B200_PAIR_TAITWATER_MULTIPHASE = 5,/* sph/taitwater/multiphase     pair_sph_taitwater_multiphase.cpp:55-186 */
  B200_PAIR_COLORGRADIENT = 6,      /* sph/colorgradient             pair_sph_colorgradient.cpp:70-191 */
  B200_PAIR_SURFACETENSION = 7,     /* sph/surfacetension            pair_sph_surfacetension.cpp:50-192 */
  B200_PAIR_HEATCONDUCTION = 8,     /* sph/heatconduction            pair_sph_heatconduction.cpp:47-134 */
  B200_PAIR_HEATCONDUCTION_MULTIPHASE = 9,  /* sph/heatconduction/multiphase  ..._multiphase.cpp:49-129 */
  B200_PAIR_HEATCONDUCTION_PHASECHANGE = 10,/* sph/heatconduction/phasechange ..._phasechange.cpp:52-141 */
  B200_PAIR_IDEALGAS = 11,          /* sph/idealgas                  pair_sph_idealgas.cpp:48-175 (coeff: I J viscosity h) */
  B200_PAIR_LJ = 12                 /* sph/lj                        pair_sph_lj.cpp:48-182 (coeff: I J viscosity h).  Its `fi += lrc` inside the
                                       neighbor loop (:139) makes every pair force depend on the reference's list order: the engine ranks
                                       each row's entries in that order (csrc/b200_lj.cuh); needs a full-list sub-style (sph/rhosum) in the deck */
};

/* One sub-style of `pair_style hybrid/overlay` (or the single pair style),
 * described by the tables the reference objects hold after Pair::init()
 * (src/pair.cpp:174-235, src/pair_hybrid.cpp:407-543).  Unused pointers NULL.
 *   mapped[i][j] = 1 iff this sub-style computes type pair (i,j)
 *                  (= !ijskip of its neighbor request, pair_hybrid.cpp:452-471)
 *   cut, cutsq   = PairSPH*::cut / Pair::cutsq of the sub-style
 * per-type:  rho0, B, soundspeed (taitwater*), gamma, rbackground (multiphase)
 * per-pair:  viscosity (taitwater*), alpha (colorgradient alpha / heat D),
 *            tc + fixflag (heatconduction/phasechange :124-129)
 * Which atom of a pair holds it in the reference's half list follows from the set of sub-styles: with a full-list style among them
 * (sph/rhosum, sph/rhosum/multiphase, sph/colorgradient: pair_sph_rhosum.cpp:60-61) the half lists are derived from the full one
 * (half_from_full_newton, neigh_derive.cpp:83-145), without one they are built by half_bin_newton (neigh_half_bin.cpp:285-420).  The
 * engine applies the matching rule wherever a result depends on the holder (stale ghost fields at setup, gamma of the list owner,
 * the phase-change clamp, the viscosity table of sph/idealgas).                */
typedef struct {
  int style;                 /* B200_PAIR_* */
  int nstep;                 /* rhosum*, colorgradient: settings() arg; else 0 */
  const int    *mapped;
  const double *cut;
  const double *cutsq;
  const double *rho0;
  const double *B;
  const double *soundspeed;
  const double *gamma;
  const double *rbackground;
  const double *viscosity;
  const double *alpha;
  const double *tc;
  const int    *fixflag;
} b200_pair_desc;

/* fix phase_change arguments, fix_phase_change.cpp:57-79 (constructor order) */
typedef struct {
  int    groupbit;
  double Tc, Tt, Hwv, dr, to_mass, cutoff;
  int    from_type, to_type, nfreq, seed;
  int    energy_chance_flag;     /* 1: "ENERGY rate" form                     */
  double change_chance;          /* prob form                                 */
  double phase_change_rate;      /* ENERGY form                               */
  int    maxattempt;             /* "attempt N", default 10 (:364)            */
  long long first_step;          /* next_reneighbor at creation = ntimestep+1 */
} b200_phase_change_desc;

/* per-atom field bundle, LAMMPS local order (atom->x, v, vest, ... src/atom.h:76-126).
 * Any pointer may be NULL (field skipped).  Arrays hold n atoms.               */
typedef struct {
  double *x, *v, *vest, *f;        /* [n][3] */
  double *rho, *drho, *e, *de, *cv, *rmass;
  double *colorgradient;           /* [n][3] */
  int    *type, *mask, *tag;
} b200_atoms;

/* ---- life cycle ---------------------------------------------------------- */
int  b200_create(b200_sph **h, int device);
int  b200_destroy(b200_sph *h);
const char *b200_last_error(void);
const char *b200_version(void);

/* ---- multi-GPU: one engine instance per rank, LAMMPS-style brick decomposition ----
 * (src/comm_brick.cpp, src/procmap.cpp).  Rank 0 creates an id, the host layer broadcasts it (MPI in
 * LAMMPS, torch.distributed in the Python driver), every rank calls b200_comm_init BEFORE b200_domain.
 * procgrid = comm->procgrid, myloc = comm->myloc, procneigh[2*dim+dir] = comm->procneigh[dim][dir].
 * Ghost exchange, reverse accumulation and atom migration then run over NCCL send/recv (NVLink). */
int  b200_comm_unique_id(char id[128]);
int  b200_comm_init(b200_sph *h, int world, int rank, const int procgrid[3], const int myloc[3], const int procneigh[6],
                    const char id[128]);

/* ---- problem definition (host state LAMMPS already parsed) ---------------- */
/* Domain: dimension, box, periodicity  (src/domain.h boxlo/boxhi/periodicity).
 * sublo/subhi = this rank's sub-domain (== box on one GPU), src/domain.h sublo. */
int  b200_domain(b200_sph *h, int dim, const double boxlo[3], const double boxhi[3],
                 const int periodicity[3], const double sublo[3], const double subhi[3]);
/* boundary command styles per face, boundary[2*dim+side] = 0 p, 1 f, 2 s, 3 m (src/domain.h:32, Domain::set_boundary
 * domain.cpp:1440-1492).  Optional: only decks with a shrink-wrapped face (s or m) need it.  The engine then re-fits the box to the
 * owned atoms at setup and on every rebuild, exactly as Domain::reset_box (domain.cpp:338-406) + comm->setup + setup_bins
 * (verlet.cpp:102, 244-248) do.  small = Domain::small (domain.cpp:184-186: 1e-4 x the box lengths when the box was created),
 * minbox[2*dim+side] = minxlo, minxhi, minylo, ... (the box given by the deck, the floor of an m face, domain.cpp:192-207).
 * Call after b200_domain; with several ranks the engine also re-derives sublo/subhi (Domain::set_local_box, domain.cpp:301-330,
 * uniform xsplit).  b200_get_box returns the current box (for thermo volume and dumps). */
int  b200_boundary(b200_sph *h, const int boundary[6], const double small[3], const double minbox[6]);
int  b200_get_box(b200_sph *h, double boxlo[3], double boxhi[3]);
/* atom_style meso (multiphase=0, atom_vec_meso.cpp) or meso/multiphase (=1);
 * mass = atom->mass [ntypes+1] (may be NULL for multiphase: rmass is used).    */
int  b200_atom_style(b200_sph *h, int multiphase, int ntypes, const double *mass);
/* neighbor <skin> bin + neigh_modify every/delay/check (neighbor.cpp:1332-1347);
 * cutneighsq = Neighbor::cutneighsq [(ntypes+1)^2] verbatim (neighbor.cpp:259-268),
 * cutneighmax likewise; cutghost = comm->cutghost (comm_brick.cpp:166-172).    */
int  b200_neighbor(b200_sph *h, double skin, int every, int delay, int check,
                   const double *cutneighsq, double cutneighmax, double cutghost);
int  b200_timestep(b200_sph *h, double dt, double ftm2v, long long ntimestep);
/* comm_modify vel yes: ghosts carry v (needed by fix phase_change)             */
int  b200_comm_modify(b200_sph *h, int ghost_velocity);
/* atom_modify sort Nfreq binsize (Atom::modify_params, src/atom.cpp:540-552; defaults 1000, 0.0 = half the neighbor cutoff).  The engine
 * keeps its own device order; Atom::sort (atom.cpp:1555-1650, called by Verlet::setup and on the first rebuild at or after nextsort,
 * verlet.cpp:106,251) is reproduced as a re-numbering of the local indices: it decides half-list ownership, the draw order of
 * fix phase_change and the order of the arrays b200_get_atoms returns.  sortfreq 0 = atom_modify sort 0 0 (never). */
int  b200_atom_modify(b200_sph *h, int sortfreq, double userbinsize);

/* Pair sub-styles in deck order (PairHybrid::compute order, pair_hybrid.cpp:101-109). */
int  b200_pair_clear(b200_sph *h);
int  b200_pair_add(b200_sph *h, const b200_pair_desc *d);      /* returns slot >= 0 */

/* Fixes, in deck order (Modify hook order, src/modify.cpp).                    */
int  b200_fix_clear(b200_sph *h);                               /* forget the registered fixes; a fix phase_change registered again with the same arguments keeps its next step and RNG position (fix_phase_change.cpp:116,345) */
int  b200_fix_meso(b200_sph *h, int groupbit);                  /* fix_meso.cpp:91-180 */
int  b200_fix_meso_stationary(b200_sph *h, int groupbit);       /* fix_meso_stationary.cpp:71-112 */
int  b200_fix_gravity(b200_sph *h, int groupbit, double xacc, double yacc, double zacc); /* fix_gravity.cpp:244-295 */
int  b200_fix_phase_change(b200_sph *h, const b200_phase_change_desc *d); /* fix_phase_change.cpp:167-352 */
/* fix setmeso (src/USER-SPH/fix_setmeso.cpp:180-271), constant form: which = 0 meso_rho | 1 meso_e | 2 meso_t (e = cv*T);
 * region_kind 0 none | 1 block (xlo xhi ylo yhi zlo zhi) | 2 sphere (xc yc zc radius); match_inside = 1 for `region`,
 * 0 for `noregion` (apply outside).  Region tests as RegBlock/RegSphere::inside (region_block.cpp:114-119, region_sphere.cpp:96-105). */
int  b200_fix_setmeso(b200_sph *h, int groupbit, int which, double value, int region_kind, const double region[6], int match_inside);
/* fix setmeso with a variable value (`v_name`, fix_setmeso.cpp:100-140,238-262): `formula` is the text of the equal- or atom-style
 * variable (src/variable.cpp), compiled for the device: numbers, PI, the atom vectors id mass type x y z vx vy vz fx fy fz, the thermo
 * keywords step and dt, the operators + - * / % ^ == != < <= > >= && || ! with the reference's precedence and left-to-right
 * association (variable.cpp:99-107,1641), sqrt exp ln log abs sin cos tan asin acos atan atan2 ceil floor round.  Anything else is
 * refused with a message. */
int  b200_fix_setmeso_var(b200_sph *h, int groupbit, int which, const char *formula, int region_kind, const double region[6], int match_inside);
/* fix addforce fx fy fz (src/fix_addforce.cpp:40-150,243-330): per component either the constant value[d] (formula[d] == NULL) or the
 * formula of the equal- / atom-style variable the deck names with v_name (the shipped body-force decks: examples/USER/sph/poiseuille/
 * poiseuille.lmp:57-58, flow_around_cylinder/flow.lmp:84-85, bubble_on_wall/bubble.lmp:187-188).  All three components are evaluated
 * on the forces as they stand before this fix, then added.  The keywords `every N` and `region ID` have no argument of their own: the
 * caller writes them as factors of the formula, `(value)*((step%N)==0)*(<the region's inside test on x y z>)` (FixAddForceB200 in
 * lammps/USER-B200/fix_b200.cpp; fix gravity with equal-style variables arrives the same way, `mass*((magnitude)*(direction))`).
 * `energy` has no counterpart. */
int  b200_fix_addforce(b200_sph *h, int groupbit, const double value[3], const char *const formula[3]);
/* host-side check of a formula (what `variable` + `fix ... v_name` would hand the two calls above): compiles it and, when atom != NULL,
 * evaluates the compiled program on the host for one atom {x y z vx vy vz fx fy fz mass}.  -1 + b200_last_error() if the formula uses
 * anything the device evaluator does not have.  Needs no device. */
int  b200_formula_check(const char *formula, const double atom[12], int type, int id, double step, double dt, double *value);
int  b200_fix_enforce2d(b200_sph *h, int groupbit);                          /* fix_enforce2d.cpp:77-89 */
/* fix setmesode value [region ID] (fix_setmesode.cpp:38-78,171-199): de = value for the group's atoms (inside the region); constant value,
 * region as for b200_fix_setmeso */
int  b200_fix_setmesode(b200_sph *h, int groupbit, double value, int region_kind, const double region[6]);
/* fix dt/reset N Tmin Tmax Xmax units box (fix_dt_reset.cpp:40-186): every N steps (and at setup) the timestep becomes the largest one
 * that moves no atom of the group further than xmax, clamped to [tmin, tmax] where minbound / maxbound != 0.  The timestep then lives on
 * the device; b200_get_timestep returns the current value (update->dt). */
int  b200_fix_dt_reset(b200_sph *h, int groupbit, int nevery, int minbound, double tmin, int maxbound, double tmax, double xmax);
int  b200_get_timestep(b200_sph *h, double *dt);
/* Elapsed simulation time under a varying timestep.  Update::update_time (update.cpp:480-484) adds (ntimestep - atimestep) * dt to atime
 * and moves atimestep whenever FixDtReset::end_of_step changes dt, and the fix remembers that step as `laststep` (fix_dt_reset.cpp:175-181;
 * thermo keyword `time` = atime + (ntimestep - atimestep) * dt, thermo.cpp:1500; f_ID of the fix = laststep, fix_dt_reset.cpp:190-193).
 * With fix dt/reset registered the engine keeps the three on the device next to dt: b200_set_time hands over the caller's values before
 * b200_setup, b200_get_time returns them after b200_setup / b200_run.  Without the fix they come back unchanged. */
int  b200_set_time(b200_sph *h, double atime, long long atimestep, long long laststep);
int  b200_get_time(b200_sph *h, double *atime, long long *atimestep, long long *laststep);
/* Pair virial (Pair::virial_fdotr_compute, pair.cpp:1403-1451: sum of x (x) f over owned + ghost atoms of the pair forces, before the
 * reverse halo; xx yy zz xy xz yz, this rank's share).  b200_request_virial arms it for the next force evaluation that ends a call:
 * the one of b200_setup, or the LAST step of the next b200_run (Verlet's ev_set on thermo steps, integrate.cpp:120-150). */
int  b200_request_virial(b200_sph *h);
int  b200_get_virial(b200_sph *h, double v[6]);
/* fix setforce with constant values (fix_setforce.cpp:215-251): set[d] != 0 -> f[d] = value[d] (set[d] = 0 is the NULL keyword) */
int  b200_fix_setforce(b200_sph *h, int groupbit, const int set[3], const double value[3]);

/* ---- per-atom data -------------------------------------------------------- */
/* Upload nlocal owned atoms (LAMMPS local order).  x,v,rho,e,type,mask,tag are
 * required; vest defaults to v, cv to 0, rmass to mass[type], cg to 0.         */
int  b200_set_atoms(b200_sph *h, int nlocal, const b200_atoms *a);
int  b200_get_natoms(b200_sph *h, int *nlocal, int *nghost);
/* Download owned atoms back in LAMMPS local order (non-NULL fields only).
 * Atoms created by fix phase_change are appended in creation order.  With more than one rank
 * the order is this rank's local order (migrated atoms appended on arrival); match ranks by tag. */
int  b200_get_atoms(b200_sph *h, int nmax, b200_atoms *a);

/* ---- the hot path --------------------------------------------------------- */
/* Verlet::setup (verlet.cpp:88-142): pbc, ghosts, neighbor build, forces.      */
int  b200_setup(b200_sph *h);
/* Verlet::run(n) (verlet.cpp:207-309), device resident, no host copies.        */
int  b200_run(b200_sph *h, int nsteps);

/* Stage-level entry points for Pair::compute / Fix hooks driven by LAMMPS' own
 * Verlet loop, and for the parity tests.                                      */
int  b200_initial_integrate(b200_sph *h);
int  b200_final_integrate(b200_sph *h);
int  b200_neigh_decide(b200_sph *h, int *rebuild);   /* Neighbor::decide                */
int  b200_forward_comm(b200_sph *h);                 /* CommBrick::forward_comm :444    */
int  b200_reneighbor(b200_sph *h);                   /* pre_exchange..neighbor->build   */
int  b200_force_clear(b200_sph *h);                  /* Verlet::force_clear :325        */
int  b200_pair_compute(b200_sph *h, int slot);       /* one PairSPH*::compute           */
int  b200_pair_compute_all(b200_sph *h);             /* PairHybrid::compute (fused)     */
int  b200_reverse_comm(b200_sph *h);                 /* CommBrick::reverse_comm :513    */
int  b200_post_force(b200_sph *h);

/* Full neighbor list of the last build (Neighbor::full_bin, neigh_full.cpp:241-340)
 * for the bit-exact check: per owned atom (LAMMPS local order) numneigh[i], then
 * entries as (tag of j, image code of j) with image = (px+1) + 3*(py+1) + 9*(pz+1),
 * 13 = owned atom; rows sorted by (tag,image).  Call with jtag==NULL to get counts. */
int  b200_get_neighbor_list(b200_sph *h, int nlocal, int *numneigh,
                            long long nentries, int *jtag, int *jimage);

/* ---- instrumentation ------------------------------------------------------ */
/* counters[0]=kernel launches, [1]=neighbor builds, [2]=steps, [3]=max neighbors,
 * [4]=nghost, [5]=row stride, [6]=phase-change insertions, [7]=dangerous builds */
int  b200_get_counters(b200_sph *h, long long counters[8]);
/* Per-stage CUDA-event timing (ms) accumulated since the last reset; names via
 * b200_timer_name(i).  Enabled with b200_set_timing(h,1) (adds event records). */
int  b200_set_timing(b200_sph *h, int on);
int  b200_get_timers(b200_sph *h, int n, double *ms, long long *calls);
const char *b200_timer_name(int i);
int  b200_sync(b200_sph *h);

#ifdef __cplusplus
}
#endif
#endif /* B200_SPH_H */
