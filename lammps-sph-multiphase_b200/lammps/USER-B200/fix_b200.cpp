#include "string.h"
#include "stdlib.h"
#include "fix_b200.h"
#include "atom.h"
#include "domain.h"
#include "error.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

void LAMMPS_NS::b200_fix_guard(LAMMPS *lmp, const char *name)
{
  if (strcmp(lmp->update->integrate_style, "verlet/b200") != 0) {
    char msg[256];
    sprintf(msg, "fix %s/b200 needs run_style verlet/b200 (use -sf b200); there is no CPU fallback", name);
    lmp->error->all(FLERR, msg);
  }
}

int FixGravityB200::b200_register(b200_sph *h)
{
  if (varflag != 0) error->all(FLERR, "fix gravity/b200 supports constant gravity only");
  set_acceleration();
  return b200_fix_gravity(h, groupbit, xacc, yacc, zacc);
}

FixPhaseChangeB200::FixPhaseChangeB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  if (narg < 14) error->all(FLERR, "Illegal fix phase_change command");
  memset(&d, 0, sizeof d);
  int m = 3;
  d.groupbit = groupbit;
  d.Tc = atof(arg[m++]); d.Tt = atof(arg[m++]); d.Hwv = atof(arg[m++]); d.dr = atof(arg[m++]);
  d.to_mass = atof(arg[m++]); d.cutoff = atof(arg[m++]);
  d.from_type = atoi(arg[m++]); d.to_type = atoi(arg[m++]); d.nfreq = atoi(arg[m++]); d.seed = atoi(arg[m++]);
  if (d.seed <= 0) error->all(FLERR, "Illegal value for seed");
  if (strcmp(arg[m++], "ENERGY") == 0) { d.energy_chance_flag = 1; d.phase_change_rate = atof(arg[m++]); }
  else { d.change_chance = atof(arg[m - 1]); if (d.change_chance < 0) error->all(FLERR, "Illegal value for change_chance"); }
  d.maxattempt = 10;
  int iregion = -1;
  while (m < narg) {
    if (strcmp(arg[m], "region") == 0) {
      if (m + 2 > narg) error->all(FLERR, "Illegal fix phase_change command");
      iregion = domain->find_region(arg[m + 1]);
      if (iregion == -1) error->all(FLERR, "Region ID for fix phase_change does not exist");
      m += 2;
    } else if (strcmp(arg[m], "attempt") == 0) {
      if (m + 2 > narg) error->all(FLERR, "Illegal fix phase_change command");
      d.maxattempt = atoi(arg[m + 1]); m += 2;
    } else if (strcmp(arg[m], "units") == 0) {
      if (m + 2 > narg || strcmp(arg[m + 1], "box") != 0) error->all(FLERR, "Illegal fix phase_change command");
      m += 2;
    } else error->all(FLERR, "Illegal fix phase_change command");
  }
  if (iregion == -1) error->all(FLERR, "Must specify a region in fix phase_change");
  force_reneighbor = 1;
  next_reneighbor = update->ntimestep + 1;
  d.first_step = next_reneighbor;
}

int FixPhaseChangeB200::setmask() { return PRE_EXCHANGE; }
int FixPhaseChangeB200::b200_register(b200_sph *h) { return b200_fix_phase_change(h, &d); }
