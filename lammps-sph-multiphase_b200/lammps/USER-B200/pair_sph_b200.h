/* -*- c++ -*- ----------------------------------------------------------
   USER-B200 pair shells: one class per USER-SPH pair style, same style name + /b200,
   same settings()/coeff() syntax (inherited from the reference class).
------------------------------------------------------------------------- */
#ifdef PAIR_CLASS

PairStyle(sph/rhosum/b200,PairSPHRhoSumB200)
PairStyle(sph/rhosum/multiphase/b200,PairSPHRhoSumMultiphaseB200)
PairStyle(sph/taitwater/b200,PairSPHTaitwaterB200)
PairStyle(sph/taitwater/morris/b200,PairSPHTaitwaterMorrisB200)
PairStyle(sph/taitwater/multiphase/b200,PairSPHTaitwaterMultiphaseB200)
PairStyle(sph/colorgradient/b200,PairSPHColorGradientB200)
PairStyle(sph/surfacetension/b200,PairSPHSurfaceTensionB200)
PairStyle(sph/heatconduction/b200,PairSPHHeatConductionB200)
PairStyle(sph/heatconduction/multiphase/b200,PairSPHHeatConductionMultiPhaseB200)
PairStyle(sph/heatconduction/phasechange/b200,PairSPHHeatConductionPhaseChangeB200)
PairStyle(sph/idealgas/b200,PairSPHIdealGasB200)
PairStyle(sph/lj/b200,PairSPHLJB200)

#else

#ifndef LMP_PAIR_SPH_B200_H
#define LMP_PAIR_SPH_B200_H

#include "b200_shell.h"
#include "pair_sph_rhosum.h"
#include "pair_sph_rhosum_multiphase.h"
#include "pair_sph_taitwater.h"
#include "pair_sph_taitwater_morris.h"
#include "pair_sph_taitwater_multiphase.h"
#include "pair_sph_colorgradient.h"
#include "pair_sph_surfacetension.h"
#include "pair_sph_heatconduction.h"
#include "pair_sph_heatconduction_multiphase.h"
#include "pair_sph_heatconduction_phasechange.h"
#include "pair_sph_idealgas.h"
#include "pair_sph_lj.h"

namespace LAMMPS_NS {

void b200_pair_compute_guard(class LAMMPS *, const char *);

#define B200_PAIR_SHELL(Class, Base)                                                        \
  class Class : public Base, public B200PairShell {                                         \
   public:                                                                                  \
    Class(class LAMMPS *lmp) : Base(lmp) { suffix_flag |= 0; }                              \
    /* the engine builds its own cell-sorted rows: no host neighbor list is requested */    \
    void init_style() {}                                                                    \
    /* Pair::compute is never the compute path: run_style verlet/b200 drives the engine */  \
    void compute(int, int) { b200_pair_compute_guard(lmp, #Base); }                         \
    void b200_describe(b200_pair_desc &, std::vector<std::vector<double> > &,               \
                       std::vector<std::vector<int> > &);                                   \
  };

B200_PAIR_SHELL(PairSPHRhoSumB200, PairSPHRhoSum)
B200_PAIR_SHELL(PairSPHRhoSumMultiphaseB200, PairSPHRhoSumMultiphase)
B200_PAIR_SHELL(PairSPHTaitwaterB200, PairSPHTaitwater)
B200_PAIR_SHELL(PairSPHTaitwaterMorrisB200, PairSPHTaitwaterMorris)
B200_PAIR_SHELL(PairSPHTaitwaterMultiphaseB200, PairSPHTaitwaterMultiphase)
B200_PAIR_SHELL(PairSPHColorGradientB200, PairSPHColorGradient)
B200_PAIR_SHELL(PairSPHSurfaceTensionB200, PairSPHSurfaceTension)
B200_PAIR_SHELL(PairSPHHeatConductionB200, PairSPHHeatConduction)
B200_PAIR_SHELL(PairSPHHeatConductionMultiPhaseB200, PairSPHHeatConductionMultiPhase)
B200_PAIR_SHELL(PairSPHIdealGasB200, PairSPHIdealGas)
B200_PAIR_SHELL(PairSPHLJB200, PairSPHLJ)

// heatconduction/phasechange leaves tc/fixflag uninitialised for the 4-argument coeff form
// (pair_sph_heatconduction_phasechange.cpp:191-218); the shell zeroes them at allocation.
class PairSPHHeatConductionPhaseChangeB200 : public PairSPHHeatConductionPhaseChange, public B200PairShell {
 public:
  PairSPHHeatConductionPhaseChangeB200(class LAMMPS *lmp) : PairSPHHeatConductionPhaseChange(lmp) {}
  void init_style() {}
  void compute(int, int) { b200_pair_compute_guard(lmp, "PairSPHHeatConductionPhaseChange"); }
  void coeff(int, char **);
  void b200_describe(b200_pair_desc &, std::vector<std::vector<double> > &, std::vector<std::vector<int> > &);
};

}    // namespace LAMMPS_NS
#endif
#endif
