#include "string.h"
#include "stdlib.h"
#include "fix_b200.h"
#include "atom.h"
#include "domain.h"
#include "region.h"
#include "region_sphere.h"
#include "error.h"
#include "update.h"
#include "input.h"
#include "variable.h"
#include <string>

using namespace LAMMPS_NS;
using namespace FixConst;

// Variable keeps the formula text of its variables private (variable.h:55 `char ***data`) and retrieve() returns NULL for atom-style
// ones (variable.cpp:647).  An explicit template instantiation may name a private member, which gives the shell a read-only view
// without touching the reference's sources.
namespace {
template <typename Tag, typename Tag::type M> struct PrivateMember { friend typename Tag::type b200_get(Tag) { return M; } };
struct VariableData { typedef char ***Variable::*type; friend type b200_get(VariableData); };
template struct PrivateMember<VariableData, &Variable::data>;
// RegSphere keeps centre and radius private (region_sphere.h:37-39).  The bounding box would give them back only up to rounding
// ((xc + r) - (xc - r) is not 2 r in floating point), and RegSphere::inside tests `r <= radius` exactly: an atom ON the sphere must not flip
struct SphereXc { typedef double RegSphere::*type; friend type b200_get(SphereXc); };
struct SphereYc { typedef double RegSphere::*type; friend type b200_get(SphereYc); };
struct SphereZc { typedef double RegSphere::*type; friend type b200_get(SphereZc); };
struct SphereR { typedef double RegSphere::*type; friend type b200_get(SphereR); };
template struct PrivateMember<SphereXc, &RegSphere::xc>;
template struct PrivateMember<SphereYc, &RegSphere::yc>;
template struct PrivateMember<SphereZc, &RegSphere::zc>;
template struct PrivateMember<SphereR, &RegSphere::radius>;
void sphere_geometry(Region *reg, double *c, double *rad)
{
  RegSphere *s = dynamic_cast<RegSphere *>(reg);
  c[0] = s->*b200_get(SphereXc()); c[1] = s->*b200_get(SphereYc()); c[2] = s->*b200_get(SphereZc()); *rad = s->*b200_get(SphereR());
}

std::string formula_of(LAMMPS *lmp, const char *name, int depth)
{
  Variable *var = lmp->input->variable;
  int ivar = var->find((char *) name);
  if (ivar < 0) lmp->error->all(FLERR, "Variable name for fix /b200 does not exist");
  if (!var->equalstyle(ivar) && !var->atomstyle(ivar)) lmp->error->all(FLERR, "Variable for fix /b200 is invalid style");
  if (depth > 8) lmp->error->all(FLERR, "Variable has circular dependency");
  char ***data = var->*b200_get(VariableData());
  std::string text = data[ivar][0], out;
  for (size_t i = 0; i < text.size();) {
    const bool start = i == 0 || !(isalnum((unsigned char) text[i - 1]) || text[i - 1] == '_');
    if (start && text.compare(i, 2, "v_") == 0) {
      size_t j = i + 2;
      while (j < text.size() && (isalnum((unsigned char) text[j]) || text[j] == '_')) j++;
      out += "(" + formula_of(lmp, text.substr(i + 2, j - i - 2).c_str(), depth + 1) + ")";
      i = j;
    } else out += text[i++];
  }
  return out;
}
}

char *LAMMPS_NS::b200_variable_formula(LAMMPS *lmp, const char *vname)
{
  std::string f = formula_of(lmp, vname, 0);
  char *s = new char[f.size() + 1];
  strcpy(s, f.c_str());
  return s;
}

void LAMMPS_NS::b200_fix_guard(LAMMPS *lmp, const char *name)
{
  if (strcmp(lmp->update->integrate_style, "verlet/b200") != 0) {
    char msg[256];
    sprintf(msg, "fix %s/b200 needs run_style verlet/b200 (use -sf b200); there is no CPU fallback", name);
    lmp->error->all(FLERR, msg);
  }
}

// Constant gravity is the engine's own fix (b200_fix_gravity).  With equal-style variables (fix_gravity.cpp:244-258: magnitude and, for
// `vector`, the direction components are re-evaluated every step) the force is handed over as the formula FixGravity::post_force
// evaluates, f += massone * (magnitude * xgrav) with xgrav = xdir / sqrt(xdir xdir + ydir ydir [+ zdir zdir]) (set_acceleration,
// :304-337), in the same association, to the engine's per-atom formula evaluator (b200_fix_addforce; csrc/b200_expr.cuh: step, dt,
// time, arithmetic and math functions -- anything else is refused there).  Variable chute / spherical angles are refused.
int FixGravityB200::b200_register(b200_sph *h)
{
  if (varflag == 0) {      // enum{CONSTANT,EQUAL}, fix_gravity.cpp:34
    set_acceleration();
    return b200_fix_gravity(h, groupbit, xacc, yacc, zacc);
  }
  if (vstyle || pstyle || tstyle) error->all(FLERR, "fix gravity/b200 supports variables for the magnitude and the components of `vector`");
  char num[64];
  std::string M, D[3];
  if (mstyle) { char *s = b200_variable_formula(lmp, mstr); M = std::string("(") + s + ")"; delete [] s; }
  else { sprintf(num, "(%.17g)", magnitude); M = num; }
  const int dim = domain->dimension;
  if (style == 2) {        // enum{CHUTE,SPHERICAL,VECTOR}: constant direction computed once, as set_acceleration does
    char *vs[3] = {xstr, ystr, zstr}; int st[3] = {xstyle, ystyle, zstyle}; double dv[3] = {xdir, ydir, zdir};
    if (!xstyle && !ystyle && !zstyle) {
      double m0 = magnitude; magnitude = 1.0; set_acceleration(); magnitude = m0;      // xacc = 1.0 * xgrav
      double g[3] = {xacc, yacc, zacc};
      for (int d = 0; d < 3; d++) { sprintf(num, "(%.17g)", g[d]); D[d] = num; }
    } else {
      std::string c[3];
      for (int d = 0; d < 3; d++) {
        if (st[d]) { char *s = b200_variable_formula(lmp, vs[d]); c[d] = std::string("(") + s + ")"; delete [] s; }
        else { sprintf(num, "(%.17g)", dv[d]); c[d] = num; }
      }
      std::string len = "sqrt(" + c[0] + "*" + c[0] + "+" + c[1] + "*" + c[1] + (dim == 3 ? "+" + c[2] + "*" + c[2] : std::string("")) + ")";
      for (int d = 0; d < 3; d++) D[d] = (d == 2 && dim == 2) ? std::string("(0.0)") : "(" + c[d] + "/" + len + ")";
    }
  } else {                 // chute / spherical with constant angles
    double m0 = magnitude; magnitude = 1.0; set_acceleration(); magnitude = m0;
    double g[3] = {xacc, yacc, zacc};
    for (int d = 0; d < 3; d++) { sprintf(num, "(%.17g)", g[d]); D[d] = num; }
  }
  std::string F[3];
  const char *f[3];
  const double zero[3] = {0.0, 0.0, 0.0};
  for (int d = 0; d < 3; d++) { F[d] = "mass*(" + M + "*" + D[d] + ")"; f[d] = F[d].c_str(); }
  return b200_fix_addforce(h, groupbit, zero, f);
}

// FixGravity::post_force sums the potential energy of the group while it adds the force (fix_gravity.cpp:262-283:
// egrav -= massone * (xacc x + yacc y + zacc z), owned atoms in index order) and compute_scalar all-reduces it (:342-351).
// The engine adds the force; the sum is formed here from the host arrays, which VerletB200 refreshes on every step that
// evaluates thermo output or an END_OF_STEP fix -- the positions are those post_force saw on that step.
double FixGravityB200::compute_scalar()
{
  double **x = atom->x;
  double *rmass = atom->rmass, *mass = atom->mass;
  int *mask = atom->mask, *type = atom->type;
  int nlocal = atom->nlocal;
  if (varflag != 0) {      // the acceleration of this step (fix_gravity.cpp:248-258)
    if (mstyle) magnitude = input->variable->compute_equal(mvar);
    if (xstyle) xdir = input->variable->compute_equal(xvar);
    if (ystyle) ydir = input->variable->compute_equal(yvar);
    if (zstyle) zdir = input->variable->compute_equal(zvar);
    set_acceleration();
  }
  egrav = 0.0;
  for (int i = 0; i < nlocal; i++)
    if (mask[i] & groupbit) {
      double massone = rmass ? rmass[i] : mass[type[i]];
      egrav -= massone * (xacc*x[i][0] + yacc*x[i][1] + zacc*x[i][2]);
    }
  MPI_Allreduce(&egrav, &egrav_all, 1, MPI_DOUBLE, MPI_SUM, world);
  return egrav_all;
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

FixSetMesoB200::FixSetMesoB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), idregion(NULL), vname(NULL)
{
  if (narg < 5) error->all(FLERR, "Illegal fix setmeso command");
  if (strcmp(arg[3], "meso_rho") == 0) which = 0;
  else if (strcmp(arg[3], "meso_e") == 0) which = 1;
  else if (strcmp(arg[3], "meso_t") == 0) which = 2;
  else error->all(FLERR, "Illegal fix setmeso command, meso_rho or meso_e must be given");
  value = 0.0;
  if (strstr(arg[4], "v_") == arg[4]) { vname = new char[strlen(arg[4]) - 1]; strcpy(vname, arg[4] + 2); }      // xstr, fix_setmeso.cpp:52-56
  else value = atof(arg[4]);
  regionflag = 1;
  int iarg = 5;
  while (iarg < narg) {
    if (strcmp(arg[iarg], "region") == 0 || strcmp(arg[iarg], "noregion") == 0) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal fix setmesode command");
      if (domain->find_region(arg[iarg + 1]) == -1) error->all(FLERR, "Region ID for fix setmesode does not exist");
      int n = strlen(arg[iarg + 1]) + 1;
      idregion = new char[n];
      strcpy(idregion, arg[iarg + 1]);
      if (strcmp(arg[iarg], "noregion") == 0) regionflag = 0;
      iarg += 2;
    } else error->all(FLERR, "Illegal fix setmesode command");
  }
}

int FixSetMesoB200::setmask() { return POST_FORCE; }

int FixSetMesoB200::b200_register(b200_sph *h)
{
  int kind = 0;
  double r[6] = {0, 0, 0, 0, 0, 0};
  if (idregion) {
    int ir = domain->find_region(idregion);
    if (ir == -1) error->all(FLERR, "Region ID for fix setmesode does not exist");
    Region *reg = domain->regions[ir];
    if (reg->dynamic_check() || !reg->interior) error->all(FLERR, "fix setmeso/b200 supports static regions with side in");
    // RegBlock keeps its bounds private; its bounding box (region.h extent_*) is exactly them.  RegSphere: see sphere_geometry
    if (strcmp(reg->style, "block") == 0) {
      kind = 1;
      r[0] = reg->extent_xlo; r[1] = reg->extent_xhi; r[2] = reg->extent_ylo; r[3] = reg->extent_yhi; r[4] = reg->extent_zlo; r[5] = reg->extent_zhi;
    } else if (strcmp(reg->style, "sphere") == 0) {
      kind = 2;
      sphere_geometry(reg, r, r + 3);
    } else error->all(FLERR, "fix setmeso/b200 supports block and sphere regions");
  }
  if (vname) {      // the variable branch of the reference tests `!match` whatever region / noregion said (fix_setmeso.cpp:247-249)
    char *f = b200_variable_formula(lmp, vname);
    int rc = b200_fix_setmeso_var(h, groupbit, which, f, kind, r, 1);
    delete [] f;
    return rc;
  }
  return b200_fix_setmeso(h, groupbit, which, value, kind, r, regionflag);
}

FixAddForceB200::FixAddForceB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  for (int d = 0; d < 3; d++) { vname[d] = NULL; value[d] = 0.0; }
  if (narg < 6) error->all(FLERR, "Illegal fix addforce command");
  for (int d = 0; d < 3; d++) {
    const char *a = arg[3 + d];
    if (strstr(a, "v_") == a) { vname[d] = new char[strlen(a) - 1]; strcpy(vname[d], a + 2); }
    else value[d] = atof(a);
  }
  every = 1; idregion = NULL;
  for (int iarg = 6; iarg < narg; iarg += 2) {      // fix_addforce.cpp:88-118
    if (iarg + 2 > narg) error->all(FLERR, "Illegal fix addforce command");
    if (strcmp(arg[iarg], "every") == 0) {
      every = atoi(arg[iarg + 1]);
      if (every <= 0) error->all(FLERR, "Illegal fix addforce command");
    } else if (strcmp(arg[iarg], "region") == 0) {
      if (domain->find_region(arg[iarg + 1]) == -1) error->all(FLERR, "Region ID for fix addforce does not exist");
      delete [] idregion;
      idregion = new char[strlen(arg[iarg + 1]) + 1];
      strcpy(idregion, arg[iarg + 1]);
    } else if (strcmp(arg[iarg], "energy") == 0) error->all(FLERR, "fix addforce/b200 does not take the energy keyword");
    else error->all(FLERR, "Illegal fix addforce command");
  }
}
int FixAddForceB200::setmask() { return POST_FORCE; }
// `every N` (the force is added on steps that are multiples of N, fix_addforce.cpp:246) and `region ID` (atoms the region matches, :262, :290)
// become factors of the formula the engine evaluates per atom and step: value * ((step % N) == 0) * (inside), each factor 1.0 or 0.0 as
// the reference's variable arithmetic gives them, so an atom the reference skips receives + value * 0.  Region tests as RegBlock::inside /
// RegSphere::inside write them (region_block.cpp, region_sphere.cpp); static block / sphere regions with side in, as for fix setmeso/b200.
int FixAddForceB200::b200_register(b200_sph *h)
{
  std::string gate;
  char num[512];
  if (every > 1) { sprintf(num, "*((step%%%d)==0)", every); gate += num; }
  if (idregion) {
    int ir = domain->find_region(idregion);
    if (ir == -1) error->all(FLERR, "Region ID for fix addforce does not exist");
    Region *reg = domain->regions[ir];
    if (reg->dynamic_check() || !reg->interior) error->all(FLERR, "fix addforce/b200 supports static regions with side in");
    if (strcmp(reg->style, "block") == 0) {
      sprintf(num, "*((x>=(%.17g))&&(x<=(%.17g))&&(y>=(%.17g))&&(y<=(%.17g))&&(z>=(%.17g))&&(z<=(%.17g)))", reg->extent_xlo, reg->extent_xhi,
              reg->extent_ylo, reg->extent_yhi, reg->extent_zlo, reg->extent_zhi);
    } else if (strcmp(reg->style, "sphere") == 0) {
      double c3[3], rad;
      sphere_geometry(reg, c3, &rad);
      const double xc = c3[0], yc = c3[1], zc = c3[2];
      sprintf(num, "*(sqrt((x-(%.17g))*(x-(%.17g))+(y-(%.17g))*(y-(%.17g))+(z-(%.17g))*(z-(%.17g)))<=(%.17g))", xc, xc, yc, yc, zc, zc, rad);
    } else error->all(FLERR, "fix addforce/b200 supports block and sphere regions");
    gate += num;
  }
  std::string F[3];
  const char *f[3] = {NULL, NULL, NULL};
  double v[3] = {value[0], value[1], value[2]};
  for (int d = 0; d < 3; d++) {
    if (vname[d]) { char *s = b200_variable_formula(lmp, vname[d]); F[d] = gate.empty() ? std::string(s) : "(" + std::string(s) + ")" + gate; delete [] s; f[d] = F[d].c_str(); }
    else if (!gate.empty() && value[d] != 0.0) { sprintf(num, "(%.17g)", value[d]); F[d] = num + gate; f[d] = F[d].c_str(); v[d] = 0.0; }
  }
  return b200_fix_addforce(h, groupbit, v, f);
}

FixSetForceB200::FixSetForceB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  if (narg != 6) error->all(FLERR, "Illegal fix setforce command (fix setforce/b200: constant values or NULL, no region)");
  for (int d = 0; d < 3; d++) {
    const char *a = arg[3 + d];
    if (strstr(a, "v_") == a) error->all(FLERR, "fix setforce/b200 supports constant values only");
    set[d] = strcmp(a, "NULL") != 0;
    value[d] = set[d] ? atof(a) : 0.0;
  }
}
int FixSetForceB200::setmask() { return POST_FORCE; }

FixSetMesodEB200::FixSetMesodEB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), idregion(NULL)
{
  if (narg < 4) error->all(FLERR, "Illegal fix setmesode command");
  if (strstr(arg[3], "v_") == arg[3] || strcmp(arg[3], "NULL") == 0) error->all(FLERR, "fix setmesode/b200 supports a constant value only");
  value = atof(arg[3]);
  int iarg = 4;
  while (iarg < narg) {
    if (strcmp(arg[iarg], "region") == 0) {
      if (iarg + 2 > narg) error->all(FLERR, "Illegal fix setmesode command");
      if (domain->find_region(arg[iarg + 1]) == -1) error->all(FLERR, "Region ID for fix setmesode does not exist");
      idregion = new char[strlen(arg[iarg + 1]) + 1];
      strcpy(idregion, arg[iarg + 1]);
      iarg += 2;
    } else error->all(FLERR, "Illegal fix setmesode command");
  }
}
int FixSetMesodEB200::setmask() { return POST_FORCE; }
int FixSetMesodEB200::b200_register(b200_sph *h)
{
  int kind = 0; double r[6] = {0, 0, 0, 0, 0, 0};
  if (idregion) {
    int ir = domain->find_region(idregion);
    if (ir == -1) error->all(FLERR, "Region ID for fix setmesode does not exist");
    Region *reg = domain->regions[ir];
    if (reg->dynamic_check() || !reg->interior) error->all(FLERR, "fix setmesode/b200 supports static regions with side in");
    if (strcmp(reg->style, "block") == 0) { kind = 1; r[0] = reg->extent_xlo; r[1] = reg->extent_xhi; r[2] = reg->extent_ylo; r[3] = reg->extent_yhi; r[4] = reg->extent_zlo; r[5] = reg->extent_zhi; }
    else if (strcmp(reg->style, "sphere") == 0) { kind = 2; sphere_geometry(reg, r, r + 3); }
    else error->all(FLERR, "fix setmesode/b200 supports block and sphere regions");
  }
  return b200_fix_setmesode(h, groupbit, value, kind, r);
}

FixDtResetB200::FixDtResetB200(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  if (narg < 7) error->all(FLERR, "Illegal fix dt/reset command");
  time_depend = 1;
  scalar_flag = 1; global_freq = 1; extscalar = 0; extvector = 0;      // fix_dt_reset.cpp:46-50
  laststep = update->ntimestep;                                        // :88-89
  nevery_ = atoi(arg[3]);
  minbound = maxbound = 1; tmin = tmax = 0.0;
  if (strcmp(arg[4], "NULL") == 0) minbound = 0; else tmin = atof(arg[4]);
  if (strcmp(arg[5], "NULL") == 0) maxbound = 0; else tmax = atof(arg[5]);
  xmax = atof(arg[6]);
  if (narg != 9 || strcmp(arg[7], "units") != 0 || strcmp(arg[8], "box") != 0) error->all(FLERR, "fix dt/reset/b200 needs `units box`");
  if (nevery_ <= 0 || xmax <= 0.0 || (minbound && tmin < 0.0) || (maxbound && tmax < 0.0) || (minbound && maxbound && tmin >= tmax))
    error->all(FLERR, "Illegal fix dt/reset command");
}
int FixDtResetB200::setmask() { return END_OF_STEP; }
