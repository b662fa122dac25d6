/* ----------------------------------------------------------------------
   USER-B200 pair shells: describe the parsed coefficient tables to the engine.
------------------------------------------------------------------------- */
#include "string.h"
#include "pair_sph_b200.h"
#include "atom.h"
#include "error.h"
#include "update.h"
#include "memory.h"

using namespace LAMMPS_NS;

void LAMMPS_NS::b200_pair_compute_guard(LAMMPS *lmp, const char *name)
{
  if (strcmp(lmp->update->integrate_style, "verlet/b200") != 0) {
    char msg[256];
    sprintf(msg, "%s /b200 pair styles need run_style verlet/b200 (use -sf b200); there is no CPU fallback", name);
    lmp->error->all(FLERR, msg);
  }
}

#define COMMON(STYLE, NSTEP)                                        \
  int n = atom->ntypes;                                             \
  memset(&d, 0, sizeof d);                                          \
  d.style = STYLE; d.nstep = NSTEP;                                 \
  d.mapped = b200_flat2i(is, NULL, setflag, n);                     \
  d.cut = b200_flat2(ds, cut, setflag, n);                          \
  d.cutsq = b200_flat2(ds, cutsq, setflag, n);

typedef std::vector<std::vector<double> > DS;
typedef std::vector<std::vector<int> > IS;

void PairSPHRhoSumB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is) { COMMON(B200_PAIR_RHOSUM, nstep) }
void PairSPHRhoSumMultiphaseB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is) { COMMON(B200_PAIR_RHOSUM_MULTIPHASE, nstep) }

void PairSPHTaitwaterB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_TAITWATER, 0)
  d.rho0 = b200_flat1(ds, rho0, n); d.B = b200_flat1(ds, B, n); d.soundspeed = b200_flat1(ds, soundspeed, n);
  d.viscosity = b200_flat2(ds, viscosity, setflag, n);
}
void PairSPHTaitwaterMorrisB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_TAITWATER_MORRIS, 0)
  d.rho0 = b200_flat1(ds, rho0, n); d.B = b200_flat1(ds, B, n); d.soundspeed = b200_flat1(ds, soundspeed, n);
  d.viscosity = b200_flat2(ds, viscosity, setflag, n);
}
void PairSPHTaitwaterMultiphaseB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_TAITWATER_MULTIPHASE, 0)
  d.rho0 = b200_flat1(ds, rho0, n); d.B = b200_flat1(ds, B, n); d.soundspeed = b200_flat1(ds, soundspeed, n);
  d.gamma = b200_flat1(ds, gamma, n); d.rbackground = b200_flat1(ds, rbackground, n);
  d.viscosity = b200_flat2(ds, viscosity, setflag, n);
}
void PairSPHColorGradientB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_COLORGRADIENT, nstep)
  d.alpha = b200_flat2(ds, alpha, setflag, n);
}
void PairSPHSurfaceTensionB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is) { COMMON(B200_PAIR_SURFACETENSION, 0) }
void PairSPHHeatConductionB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_HEATCONDUCTION, 0)
  d.alpha = b200_flat2(ds, alpha, setflag, n);
}
void PairSPHIdealGasB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_IDEALGAS, 0)
  d.viscosity = b200_flat2(ds, viscosity, setflag, n);
}
void PairSPHLJB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_LJ, 0)
  d.viscosity = b200_flat2(ds, viscosity, setflag, n);
}
void PairSPHHeatConductionMultiPhaseB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_HEATCONDUCTION_MULTIPHASE, 0)
  d.alpha = b200_flat2(ds, alpha, setflag, n);
}

void PairSPHHeatConductionPhaseChangeB200::coeff(int narg, char **arg)
{
  if (!allocated) {
    allocate();
    int n = atom->ntypes;
    for (int i = 0; i <= n; i++)
      for (int j = 0; j <= n; j++) { tc[i][j] = 0.0; fixflag[i][j] = 0; }
  }
  PairSPHHeatConductionPhaseChange::coeff(narg, arg);
}
void PairSPHHeatConductionPhaseChangeB200::b200_describe(b200_pair_desc &d, DS &ds, IS &is)
{
  COMMON(B200_PAIR_HEATCONDUCTION_PHASECHANGE, 0)
  d.alpha = b200_flat2(ds, alpha, setflag, n);
  d.tc = b200_flat2(ds, tc, setflag, n);
  d.fixflag = b200_flat2i(is, fixflag, setflag, n);
}
