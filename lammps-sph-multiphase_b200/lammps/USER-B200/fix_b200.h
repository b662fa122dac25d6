/* -*- c++ -*- ----------------------------------------------------------
   USER-B200 fix shells: fix meso/b200, meso/stationary/b200, gravity/b200, phase_change/b200.
   Same arguments as the reference fixes; the per-step work runs inside the engine.
------------------------------------------------------------------------- */
#ifdef FIX_CLASS

FixStyle(meso/b200,FixMesoB200)
FixStyle(meso/stationary/b200,FixMesoStationaryB200)
FixStyle(gravity/b200,FixGravityB200)
FixStyle(phase_change/b200,FixPhaseChangeB200)
FixStyle(setmeso/b200,FixSetMesoB200)
FixStyle(enforce2d/b200,FixEnforce2DB200)
FixStyle(setforce/b200,FixSetForceB200)
FixStyle(setmesode/b200,FixSetMesodEB200)
FixStyle(dt/reset/b200,FixDtResetB200)
FixStyle(addforce/b200,FixAddForceB200)

#else

#ifndef LMP_FIX_B200_H
#define LMP_FIX_B200_H

#include "b200_shell.h"
#include "fix_meso.h"
#include "fix_meso_stationary.h"
#include "fix_gravity.h"
#include "fix_enforce2d.h"

namespace LAMMPS_NS {

void b200_fix_guard(class LAMMPS *, const char *);

class FixMesoB200 : public FixMeso, public B200FixShell {
 public:
  FixMesoB200(class LAMMPS *lmp, int narg, char **arg) : FixMeso(lmp, narg, arg) {}
  void setup_pre_force(int) {}                         // vest = v happens in b200_setup (k_setup_pre_force)
  void initial_integrate(int) { b200_fix_guard(lmp, "meso"); }
  void final_integrate() { b200_fix_guard(lmp, "meso"); }
  int b200_register(b200_sph *h) { return b200_fix_meso(h, groupbit); }
};

class FixMesoStationaryB200 : public FixMesoStationary, public B200FixShell {
 public:
  FixMesoStationaryB200(class LAMMPS *lmp, int narg, char **arg) : FixMesoStationary(lmp, narg, arg) {}
  void initial_integrate(int) { b200_fix_guard(lmp, "meso/stationary"); }
  void final_integrate() { b200_fix_guard(lmp, "meso/stationary"); }
  int b200_register(b200_sph *h) { return b200_fix_meso_stationary(h, groupbit); }
};

class FixGravityB200 : public FixGravity, public B200FixShell {
 public:
  FixGravityB200(class LAMMPS *lmp, int narg, char **arg) : FixGravity(lmp, narg, arg) {}
  void setup(int) {}
  void post_force(int) { b200_fix_guard(lmp, "gravity"); }
  double compute_scalar();                             // f_ID in thermo / variables (water_collapse.lmp: v_etot = c_esph+c_ke+f_gfix)
  int b200_register(b200_sph *h);
};

// FixPhaseChange keeps its parameters private (fix_phase_change.h:40-70), so the shell parses the same
// argument list itself: fix ID grp phase_change Tc Tt Hwv dr to_mass cutoff from_type to_type nfreq seed
//                       (prob | ENERGY rate) region ID [attempt N] [units box]   (fix_phase_change.cpp:57-79,358-390)
class FixPhaseChangeB200 : public Fix, public B200FixShell {
 public:
  FixPhaseChangeB200(class LAMMPS *, int, char **);
  int setmask();
  void pre_exchange() { b200_fix_guard(lmp, "phase_change"); }
  int b200_register(b200_sph *h);
 private:
  b200_phase_change_desc d;
};

// text of an equal- or atom-style variable with its v_name references spliced in (parenthesised), for the device evaluator of the
// engine (csrc/b200_expr.cuh); errors out for other variable styles
char *b200_variable_formula(class LAMMPS *, const char *vname);

// FixSetMeso keeps its parameters private (fix_setmeso.h:42-53): same argument list re-parsed,
//   fix ID grp setmeso meso_rho|meso_e|meso_t value|v_name [region|noregion ID]      (fix_setmeso.cpp:38-89)
class FixSetMesoB200 : public Fix, public B200FixShell {
 public:
  FixSetMesoB200(class LAMMPS *, int, char **);
  ~FixSetMesoB200() { delete [] idregion; delete [] vname; }
  int setmask();
  void post_force(int) { b200_fix_guard(lmp, "setmeso"); }
  int b200_register(b200_sph *h);
 private:
  int which, regionflag;
  double value;
  char *idregion, *vname;
};

// FixAddForce (fix_addforce.cpp:40-150, members private): fix ID grp addforce fx fy fz, each a constant or v_name (equal- or atom-style
// variable, evaluated per atom and step on the device), `every N` and `region ID` (static block / sphere) as factors of that formula;
// the energy keyword is refused
class FixAddForceB200 : public Fix, public B200FixShell {
 public:
  FixAddForceB200(class LAMMPS *, int, char **);
  ~FixAddForceB200() { for (int d = 0; d < 3; d++) delete [] vname[d]; delete [] idregion; }
  int setmask();
  void post_force(int) { b200_fix_guard(lmp, "addforce"); }
  int b200_register(b200_sph *h);
 private:
  double value[3]; char *vname[3];
  int every; char *idregion;
};

class FixEnforce2DB200 : public FixEnforce2D, public B200FixShell {
 public:
  FixEnforce2DB200(class LAMMPS *lmp, int narg, char **arg) : FixEnforce2D(lmp, narg, arg) {}
  void setup(int) {}
  void post_force(int) { b200_fix_guard(lmp, "enforce2d"); }
  int b200_register(b200_sph *h) { return b200_fix_enforce2d(h, groupbit); }
};

// FixSetForce keeps its parameters private (fix_setforce.h:41-46): same argument list re-parsed,
//   fix ID grp setforce fx fy fz     (fix_setforce.cpp:40-110), constant values or NULL, no region
class FixSetForceB200 : public Fix, public B200FixShell {
 public:
  FixSetForceB200(class LAMMPS *, int, char **);
  int setmask();
  void post_force(int) { b200_fix_guard(lmp, "setforce"); }
  int b200_register(b200_sph *h) { return b200_fix_setforce(h, groupbit, set, value); }
 private:
  int set[3]; double value[3];
};

// FixSetMesodE (fix_setmesode.cpp:38-78): fix ID grp setmesode value [region ID], constant value, static block / sphere region
class FixSetMesodEB200 : public Fix, public B200FixShell {
 public:
  FixSetMesodEB200(class LAMMPS *, int, char **);
  int setmask();
  void post_force(int) { b200_fix_guard(lmp, "setmesode"); }
  int b200_register(b200_sph *h);
 private:
  double value; char *idregion;
};

// FixDtReset (fix_dt_reset.cpp:40-98, members private): fix ID grp dt/reset N Tmin Tmax Xmax units box.  The timestep lives on the
// device during a run; VerletB200 copies it back into update->dt after every b200_run segment, together with update->atime / atimestep
// (thermo keyword `time`) and `laststep` (f_ID).
class FixDtResetB200 : public Fix, public B200FixShell {
 public:
  FixDtResetB200(class LAMMPS *, int, char **);
  int setmask();
  void end_of_step() { b200_fix_guard(lmp, "dt/reset"); }
  double compute_scalar() { return (double) laststep; }      // fix_dt_reset.cpp:190-193
  int b200_register(b200_sph *h) { return b200_fix_dt_reset(h, groupbit, nevery_, minbound, tmin, maxbound, tmax, xmax); }
  bigint laststep;                                     // last step the engine changed the timestep on (VerletB200 copies it back with update->atime)
 private:
  int nevery_, minbound, maxbound; double tmin, tmax, xmax;
};

}    // namespace LAMMPS_NS
#endif
#endif
