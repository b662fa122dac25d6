/* -*- c++ -*- ----------------------------------------------------------
   USER-B200: LAMMPS-side shells of the B200 SPH engine (libb200sph.so).

   The classes of this package register the reference's own USER-SPH style
   names with a /b200 suffix (pair sph/.../b200, fix meso/b200, ...,
   run_style verlet/b200), parse `settings`/`coeff`/fix arguments exactly as
   the originals do (they derive from them), and hand the parsed tables and the
   per-atom arrays across the C-ABI of include/b200_sph.h.  No SPH arithmetic
   and no CPU fallback lives here.
------------------------------------------------------------------------- */
#ifndef LMP_B200_SHELL_H
#define LMP_B200_SHELL_H

#include <vector>
#include "b200_sph.h"

namespace LAMMPS_NS {

// a pair sub-style that can describe itself to the engine (tables after Pair::init)
class B200PairShell {
 public:
  virtual ~B200PairShell() {}
  // fills d with pointers into `store` (kept alive by the caller until b200_pair_add returns)
  virtual void b200_describe(b200_pair_desc &d, std::vector<std::vector<double> > &dstore,
                             std::vector<std::vector<int> > &istore) = 0;
};

// a fix whose per-step work runs inside the engine
class B200FixShell {
 public:
  virtual ~B200FixShell() {}
  virtual int b200_register(b200_sph *h) = 0;    // calls the matching b200_fix_* entry point
};

// helpers shared by the pair shells: flatten LAMMPS' (n+1)x(n+1) tables, uninitialised entries -> 0
inline const double *b200_flat2(std::vector<std::vector<double> > &store, double **a, int **setflag, int n)
{
  store.push_back(std::vector<double>((n + 1) * (n + 1), 0.0));
  std::vector<double> &v = store.back();
  if (a)
    for (int i = 1; i <= n; i++)
      for (int j = 1; j <= n; j++) {
        int lo = i < j ? i : j, hi = i < j ? j : i;
        if (setflag[lo][hi]) v[i * (n + 1) + j] = a[i][j];   // symmetric copies exist after init_one()
      }
  return v.data();
}
inline const int *b200_flat2i(std::vector<std::vector<int> > &store, int **a, int **setflag, int n)
{
  store.push_back(std::vector<int>((n + 1) * (n + 1), 0));
  std::vector<int> &v = store.back();
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      int lo = i < j ? i : j, hi = i < j ? j : i;
      if (setflag[lo][hi]) v[i * (n + 1) + j] = a ? a[i][j] : 1;
    }
  return v.data();
}
inline const double *b200_flat1(std::vector<std::vector<double> > &store, double *a, int n)
{
  store.push_back(std::vector<double>(n + 1, 0.0));
  std::vector<double> &v = store.back();
  if (a) for (int i = 1; i <= n; i++) v[i] = a[i];
  return v.data();
}

}    // namespace LAMMPS_NS
#endif
