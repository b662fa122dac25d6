/* -*- c++ -*- ----------------------------------------------------------
   run_style verlet/b200: the Verlet loop (src/verlet.cpp:88-309) executed device-resident by
   the engine.  Host arrays are refreshed on output steps and at the end of the run only.
------------------------------------------------------------------------- */
#ifdef INTEGRATE_CLASS

IntegrateStyle(verlet/b200,VerletB200)

#else

#ifndef LMP_VERLET_B200_H
#define LMP_VERLET_B200_H

#include "verlet.h"
#include <vector>
#include "b200_sph.h"

namespace LAMMPS_NS {

class VerletB200 : public Verlet {
 public:
  VerletB200(class LAMMPS *, int, char **);
  ~VerletB200();
  void init();
  void setup();
  void setup_minimal(int);
  void run(int);

 private:
  b200_sph *h;
  std::vector<int> host_every;   // nevery of the host-side END_OF_STEP fixes (fix print, fix ave/...): segment boundaries of run()
  class FixDtResetB200 *dtfix;   // the deck's fix dt/reset/b200, if any: update->atime / atimestep and its laststep follow the engine
  void check(int rc);
  void pull_time();
  void configure();
  void upload();
  void download();
};

}
#endif
#endif
