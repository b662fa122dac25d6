"""Parity cases: small decks that exercise every style on the hot path.

Each case has
  create : LAMMPS commands (our own text, in the reference's deck language) that
           build the box and the atoms -- executed ONLY by the real reference when
           the golden fixtures are generated (tests/golden/make_golden.py);
  cmds   : the hot-path commands as tuples, applied verbatim both to the reference
           (rendered to text by `lammps_text`) and to the host mirror `Deck`;
  nsteps : length of the trajectory check.
The reference-generated initial state travels in tests/golden/<name>.npz.
"""
import importlib
import math

pkg = importlib.import_module("lammps-sph-multiphase_b200")
Deck = pkg.Deck


def _f(v):
    return repr(float(v)) if isinstance(v, float) else str(v)


class Case:
    def __init__(self, name, dim, boundary, box, atom_style, ntypes, create, cmds, nsteps, units="si", groups=(),
                 tol_traj=1e-9, regions=(), sort=(0, 0.0)):
        self.name, self.dim, self.boundary, self.box = name, dim, boundary, box
        self.atom_style, self.ntypes, self.create, self.cmds, self.nsteps = atom_style, ntypes, create, cmds, nsteps
        self.units, self.groups, self.tol_traj, self.regions = units, groups, tol_traj, regions
        self.sort = sort           # atom_modify sort Nfreq binsize; None = the LAMMPS default (1000, half the neighbor cutoff)
        self.engine = True         # False: restated by the oracle only

    @property
    def multiphase(self):
        return self.atom_style == "meso/multiphase"

    def header_text(self):
        (x0, y0, z0), (x1, y1, z1) = self.box
        return "\n".join([
            "units %s" % self.units, "dimension %d" % self.dim, "boundary %s" % self.boundary, "newton on",
            "atom_style %s" % self.atom_style, "atom_modify map array" + ("" if self.sort is None else " sort %d %s" % (self.sort[0], _f(float(self.sort[1])))),
            "region box block %s %s %s %s %s %s units box" % tuple(_f(float(v)) for v in (x0, x1, y0, y1, z0, z1)),
            "create_box %d box" % self.ntypes])

    def lammps_text(self):
        out = []
        for g, t in self.groups:
            out.append("group %s type %d" % (g, t))
        for r in self.regions:
            out.append("region %s %s %s units box" % (r[0], r[1], " ".join(_f(v) for v in r[2:])))
        nfix = 0
        for c in self.cmds:
            k, a = c[0], c[1:]
            if k == "pair_style":
                out.append("pair_style " + " ".join(str(v) for v in a))
            elif k == "pair_coeff":
                out.append("pair_coeff " + " ".join(_f(v) for v in a))
            elif k == "mass":
                out.append("mass %s %s" % (a[0], _f(a[1])))
            elif k == "neighbor":
                out.append("neighbor %s bin" % _f(a[0]))
            elif k == "neigh_modify":
                out.append("neigh_modify " + " ".join("%s %s" % (kk, vv) for kk, vv in a[0].items()))
            elif k == "comm_modify":
                out.append("comm_modify vel %s" % a[0])
            elif k == "timestep":
                out.append("timestep %s" % _f(a[0]))
            elif k == "variable":
                out.append("variable %s %s %s" % (a[0], a[1], a[2]))
            elif k == "fix":
                nfix += 1
                out.append("fix f%d %s %s %s" % (nfix, a[0], a[1], " ".join(_f(v) for v in a[2:])))
            else:
                raise ValueError(k)
        return "\n".join(out)

    def deck(self):
        d = Deck(dimension=self.dim, boundary=self.boundary, box=self.box, atom_style=self.atom_style,
                 ntypes=self.ntypes, units=self.units)
        if self.sort is not None:
            d.atom_modify(sort=self.sort)
        for g, t in self.groups:
            d.group(g)
        for r in self.regions:
            d.region(*r)
        nfix = 0
        for c in self.cmds:
            k, a = c[0], c[1:]
            if k == "pair_style":
                d.pair_style(*a)
            elif k == "pair_coeff":
                d.pair_coeff(*a)
            elif k == "mass":
                d.mass(*a)
            elif k == "neighbor":
                d.neighbor(a[0])
            elif k == "neigh_modify":
                d.neigh_modify(**a[0])
            elif k == "comm_modify":
                d.comm_modify(a[0])
            elif k == "timestep":
                d.timestep(a[0])
            elif k == "variable":
                d.variable(*a)
            elif k == "fix":
                nfix += 1
                d.fix("f%d" % nfix, a[0], a[1], *a[2:])
        return d.init()


def _single(atoms):
    return "\n".join("create_atoms %d single %s %s %s units box" % (t, _f(float(x)), _f(float(y)), _f(float(z)))
                     for t, (x, y, z) in atoms)


KATBOX = ((0, 0, -10), (10, 10, 10))
KAT3 = [(1, (5, 5, 5)), (2, (5.5, 5, 5)), (2, (5, 5, 4.8))]
_kat_tail = [("neighbor", 0.0), ("comm_modify", "yes"), ("timestep", 0.0), ("fix", "all", "meso")]

CASES = {}


def _add(c):
    CASES[c.name] = c


# ---- the six known-answer decks of examples/USER/sph/multiphase_two_atoms ----
_add(Case("kat_rhosum_multiphase", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _single(KAT3) + "\nset type 1 mass 2\nset type 2 mass 1\nset type 1 meso_rho 1\nset type 2 meso_rho 1",
          [("pair_style", "sph/rhosum/multiphase", 1), ("pair_coeff", "* *", 1.0)] + _kat_tail, 1))
_add(Case("kat_taitwater_multiphase", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _single(KAT3) + "\nset type 1 mass 2\nset type 2 mass 1\nset type 1 meso_rho 1\nset type 2 meso_rho 1",
          [("pair_style", "sph/taitwater/multiphase"), ("pair_coeff", "* *", 1.0, 1.0, 0.0, 1.0, 1.0, 0.5)] + _kat_tail, 1))
_add(Case("kat_colorgradient", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _single(KAT3) + "\nset type 1 mass 2\nset type 2 mass 1\nset type 1 meso_rho 1\nset type 2 meso_rho 1",
          [("pair_style", "sph/colorgradient", 1), ("pair_coeff", "1 1", 1.0, 0.0), ("pair_coeff", "2 2", 1.0, 0.0),
           ("pair_coeff", "1 2", 1.0, 1.0)] + _kat_tail, 1))
_add(Case("kat_surfacetension", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _single([(1, (4.6, 5.3, 5)), (2, (5.5, 5, 5.2)), (2, (5.0, 5.0, 5.0))])
          + "\nset type 1 mass 1\nset type 2 mass 1\nset type 1 meso_rho 1\nset type 2 meso_rho 1",
          [("pair_style", "hybrid/overlay", "sph/colorgradient 1", "sph/surfacetension"),
           ("pair_coeff", "1 1", "sph/colorgradient", 1.0, 0.0), ("pair_coeff", "2 2", "sph/colorgradient", 1.0, 0.0),
           ("pair_coeff", "1 2", "sph/colorgradient", 1.0, 1.0), ("pair_coeff", "1 1", "sph/surfacetension", 1.0),
           ("pair_coeff", "2 2", "sph/surfacetension", 1.0), ("pair_coeff", "1 2", "sph/surfacetension", 1.0)] + _kat_tail, 1))
_kat_heat_atoms = (_single([(1, (5, 5, 5)), (2, (5.6, 5, 5))]) + "\nset type 1 meso_rho 1\nset type 2 meso_rho 1\n"
                   "set type 1 meso_cv 3.0\nset type 2 meso_cv 1.0\n")
_add(Case("kat_heatconduction_phasechange", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _kat_heat_atoms + "set type 1 mass 1\nset type 2 mass 2\nset type 1 meso_e 1.0\nset type 2 meso_e 2.0",
          [("pair_style", "sph/heatconduction/phasechange"), ("pair_coeff", "1 1", 1.0, 1.0), ("pair_coeff", "1 2", 1.0, 1.0),
           ("pair_coeff", "2 2", 1.0, 1.0)] + _kat_tail, 1))
_add(Case("kat_phase_change", 3, "p p p", KATBOX, "meso/multiphase", 2,
          _kat_heat_atoms + "set type 1 mass 10\nset type 2 mass 2\nset type 1 meso_e 10.0\nset type 2 meso_e 2.0",
          [("pair_style", "sph/colorgradient", 1), ("pair_coeff", "1 1", 1.0, 0.0), ("pair_coeff", "2 2", 1.0, 0.0),
           ("pair_coeff", "1 2", 1.0, 1.0), ("neighbor", 0.0), ("comm_modify", "yes"), ("timestep", 0.0),
           ("fix", "all", "phase_change", 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1, 2, 1, 123456, 1.0, "region", "box", "units", "box"),
           ("fix", "all", "meso")], 1))

# ---- C1: the shipped 2-D heat-conduction deck (heatconduction/sph_heat_conduction_2d.lmp) ----
_heat2d_create = """lattice sq 0.01
create_atoms 1 box
region left block EDGE 0.499 EDGE EDGE EDGE EDGE units box
region right block 0.5 EDGE EDGE EDGE EDGE EDGE units box
set region left meso_e 1.0
set region right meso_e 2.0
set group all meso_rho 0.1"""
_add(Case("heat2d", 2, "f p p", ((0, 0, 0), (1.0, 0.1, 0.001)), "meso", 1, _heat2d_create,
          [("mass", "1", 1.0e-5), ("pair_style", "sph/heatconduction"), ("pair_coeff", "1 1", 1.0e-4, 2.0e-2),
           ("timestep", 0.025), ("neighbor", 0.002), ("fix", "all", "meso/stationary")], 160))
# C1 variant with density summation + fix meso (covers rhosum + fix meso on the C1 geometry)
_add(Case("heat2d_rhosum", 2, "f p p", ((0, 0, 0), (1.0, 0.1, 0.001)), "meso", 1, _heat2d_create,
          [("mass", "1", 1.0e-5), ("pair_style", "hybrid/overlay", "sph/rhosum 1", "sph/heatconduction"),
           ("pair_coeff", "1 1", "sph/rhosum", 2.0e-2), ("pair_coeff", "1 1", "sph/heatconduction", 1.0e-4, 2.0e-2),
           ("timestep", 0.025), ("neighbor", 0.002), ("fix", "all", "meso")], 40))
# fix setmesode: a constant heating rate inside a block region of the C1 deck (fix_setmesode.cpp)
_add(Case("heat2d_setmesode", 2, "f p p", ((0, 0, 0), (1.0, 0.1, 0.001)), "meso", 1, _heat2d_create,
          [("mass", "1", 1.0e-5), ("pair_style", "sph/heatconduction"), ("pair_coeff", "1 1", 1.0e-4, 2.0e-2),
           ("timestep", 0.025), ("neighbor", 0.002), ("fix", "all", "meso/stationary"), ("fix", "all", "setmesode", 0.5, "region", "rheat")],
          40, regions=(("rheat", "block", 0.3, 0.7, "EDGE", "EDGE", "EDGE", "EDGE"),)))
# fix setmeso with an atom-style variable (fix_setmeso.cpp:238-262; bubble_on_wall/bubble.lmp:143-144 sets a temperature profile this way):
# the left third of the C1 bar is held on a profile in x, every step
_add(Case("heat2d_setmeso_var", 2, "f p p", ((0, 0, 0), (1.0, 0.1, 0.001)), "meso", 1, _heat2d_create,
          [("mass", "1", 1.0e-5), ("pair_style", "sph/heatconduction"), ("pair_coeff", "1 1", 1.0e-4, 2.0e-2),
           ("timestep", 0.025), ("neighbor", 0.002), ("fix", "all", "meso/stationary"),
           ("variable", "eprof", "atom", "1.0+sqrt(x)*(x<=0.3)+abs(-0.5)*(x>0.3)"), ("fix", "all", "setmeso", "meso_e", "v_eprof", "region", "rleft")],
          40, regions=(("rleft", "block", "EDGE", 0.33, "EDGE", "EDGE", "EDGE", "EDGE"),)))
_add(Case("heat3d", 3, "f p p", ((0, 0, 0), (0.4, 0.08, 0.08)), "meso", 1,
          """lattice sc 0.01
create_atoms 1 box
region left block EDGE 0.199 EDGE EDGE EDGE EDGE units box
region right block 0.2 EDGE EDGE EDGE EDGE EDGE units box
set region left meso_e 1.0
set region right meso_e 2.0
set group all meso_rho 10.0""",
          [("mass", "1", 1.0e-5), ("pair_style", "sph/heatconduction"), ("pair_coeff", "1 1", 1.0e-4, 2.0e-2),
           ("timestep", 0.025), ("neighbor", 0.002), ("neigh_modify", dict(every=20, delay=0, check="no")),
           ("fix", "all", "meso/stationary")], 30))


# ---- C2 scaled down: dam break, sph/rhosum (1 1 only) + sph/taitwater, walls, gravity ----
def _dam(name, dim, nsteps, morris=False):
    dx, h, c = 0.01, 0.03, 30.0
    if dim == 2:
        box = ((0, 0, -0.001), (0.60, 0.44, 0.001)); lat = "sq"
        water = "region water block 0.03 0.23 0.03 0.33 EDGE EDGE units box"
        inner = "region inner block 0.03 0.57 0.03 EDGE EDGE EDGE units box"
        m = 1000.0 * dx * dx; grav = ("gravity", -9.81, "vector", 0, 1, 0); bnd = "f f p"
    else:
        box = ((0, 0, 0), (0.22, 0.16, 0.16)); lat = "sc"
        water = "region water block 0.03 0.10 0.03 0.10 0.03 0.10 units box"
        inner = "region inner block 0.03 0.19 0.03 0.13 0.03 EDGE units box"
        m = 1000.0 * dx ** 3; grav = ("gravity", -9.81, "vector", 0, 0, 1); bnd = "f f f"
    # walls = lattice sites of the box outside the open-top inner region
    create = """lattice %s %s origin 0.5 0.5 %s
%s
%s
create_atoms 2 box
delete_atoms region inner
create_atoms 1 region water
set group all meso_rho 1000.0
set group all meso_e 0.0""" % (lat, _f(dx), "0.5" if dim == 3 else "0", water, inner)
    tait = "sph/taitwater/morris" if morris else "sph/taitwater"
    nu = 1.0e-3 if morris else 1.0
    cmds = [("mass", "*", m), ("pair_style", "hybrid/overlay", "sph/rhosum 1", tait),
            ("pair_coeff", "* *", tait, 1000.0, c, nu, h), ("pair_coeff", "1 1", "sph/rhosum", h),
            ("fix", "water", *grav), ("fix", "water", "meso"), ("fix", "bc", "meso/stationary"),
            ("neigh_modify", dict(every=5, delay=0, check="no")), ("neighbor", 0.3 * h), ("timestep", 0.1 * h / c)]
    return Case(name, dim, bnd, box, "meso", 2, create, cmds, nsteps, groups=(("bc", 2), ("water", 1)))


_add(_dam("dam2d", 2, 60))
# examples/USER/sph/water_collapse/water_collapse.lmp:32: the shipped 2-D dam break runs with a variable timestep
_c = _dam("dam2d_dtreset", 2, 60)
_c.cmds = _c.cmds + [("fix", "all", "dt/reset", 1, "NULL", 0.1 * 0.03 / 30.0, 2.0e-8, "units", "box")]      # xmax small enough to bite from step 0
_add(_c)
# fix addforce with atom-style variables, the body-force idiom of the shipped channel decks (poiseuille.lmp:57-58 `mass*${gx}*((y<..)-(y>..))`,
# flow_around_cylinder/flow.lmp:84-85, bubble_on_wall/bubble.lmp:187-188 `mass*${gy}`): gravity as mass * g, plus a shear force that flips sign at y = 0.2
_c = _dam("dam2d_addforce", 2, 40)
_c.cmds = [c for c in _c.cmds if not (c[0] == "fix" and c[2] == "gravity")]
_c.cmds = [("variable", "bodyfy", "atom", "mass*-9.81"), ("variable", "gx", "equal", "0.5*2.0^2/4"),
           ("variable", "bodyfx", "atom", "mass*v_gx*((y<0.2)-(y>0.2))"), ("fix", "water", "addforce", "v_bodyfx", "v_bodyfy", 0.0)] + _c.cmds
_add(_c)
_add(_dam("dam3d", 3, 25))
_add(_dam("dam2d_morris", 2, 40, morris=True))


# ---- shock tube (examples/USER/sph/shock_tube/shock{2,3}d.lmp): sph/rhosum + sph/idealgas, per-type masses, fix setforce;
#      the shipped deck shrink-wraps x (boundary s p p); here the tube is shortened and periodic in x (a second contact at the wrap)
def _shock(name, dim, nsteps, onetype=False, bnd="p p p", fill="box"):
    if dim == 3:
        box = ((-12, -4, -4), (18, 4, 4)); lat = "sc"
        right = "region right block 1 EDGE EDGE EDGE EDGE EDGE units box"
        sf = ("setforce", "NULL", 0.0, 0.0)
    else:
        box = ((-30, -4, -0.05), (45, 4, 0.05)); lat = "sq"
        right = "region right block 1 EDGE EDGE EDGE EDGE EDGE units box"
        sf = ("setforce", "NULL", 0.0, 0.0)
    if fill != "box":      # atoms start short of the low-x face (an m face keeps the box there until the gas reaches it)
        lat += " 1.0\nregion fill block %s EDGE EDGE EDGE EDGE EDGE units box" % fill
        fill = "region fill"
    else:
        lat += " 1.0"
    create = """lattice %s
create_atoms 1 %s
%s
set region right type 2
set type 1 meso_e 2.5
set type 2 meso_e 0.625
set type 1 meso_rho 1.0
set type 2 meso_rho 0.25""" % (lat, fill, right)
    if onetype:      # one atom type (symmetric coefficient tables -> the tile path), a hot and a cold half
        create = """lattice %s
create_atoms 1 %s
%s
set group all meso_e 2.5
set region right meso_e 0.625
set group all meso_rho 1.0
displace_atoms all random 0.05 0.05 %s 4711 units box""" % (lat, fill, right, "0.05" if dim == 3 else "0.0")
        cmds = [("mass", "1", 1.0), ("pair_style", "hybrid/overlay", "sph/rhosum 1", "sph/idealgas"),
                ("pair_coeff", "* *", "sph/rhosum", 4.0), ("pair_coeff", "* *", "sph/idealgas", 0.75, 4.0),
                ("neighbor", 0.5), ("neigh_modify", dict(every=5, delay=0, check="yes")), ("timestep", 0.05),
                ("fix", "all", "meso"), ("fix", "all", *sf)]
        return Case(name, dim, bnd, box, "meso", 1, create, cmds, nsteps, units="lj")
    cmds = [("mass", "1", 1.0), ("mass", "2", 0.25), ("pair_style", "hybrid/overlay", "sph/rhosum 1", "sph/idealgas"),
            ("pair_coeff", "* *", "sph/rhosum", 4.0), ("pair_coeff", "* *", "sph/idealgas", 0.75, 4.0),
            ("neighbor", 0.5), ("neigh_modify", dict(every=5, delay=0, check="yes")), ("timestep", 0.05),
            ("fix", "all", "meso"), ("fix", "all", *sf)]
    return Case(name, dim, bnd, box, "meso", 2, create, cmds, nsteps, units="lj")


_add(_shock("shock3d", 3, 20))
_add(_shock("shock2d", 2, 40))
_add(_shock("gas3d", 3, 15, onetype=True))


# ---- sph/lj: Lennard-Jones EOS fluid; list-order dependent (csrc/b200_lj.cuh) ----
def _lj(name, dim, nsteps):
    box = ((0, 0, 0), (12, 12, 12)) if dim == 3 else ((0, 0, -0.05), (20.5, 20.5, 0.05))
    create = """lattice %s %s
create_atoms 1 box
set group all meso_e 1.5
set group all meso_cv 1.0
set group all meso_rho 0.6
displace_atoms all random 0.08 0.08 %s 4711 units box
mass 1 1.0
velocity all create 0.05 4711 dist gaussian""" % ("sc" if dim == 3 else "sq", "0.3" if dim == 3 else "0.6", "0.08" if dim == 3 else "0.0")
    cmds = [("mass", "1", 1.0), ("pair_style", "hybrid/overlay", "sph/rhosum 1", "sph/lj"),
            ("pair_coeff", "* *", "sph/rhosum", 2.5), ("pair_coeff", "* *", "sph/lj", 0.5, 2.5),
            ("neighbor", 0.1 if dim == 2 else 0.3), ("neigh_modify", dict(every=2, delay=0, check="yes")), ("timestep", 0.005 if dim == 2 else 0.001),
            ("fix", "all", "meso")]
    return Case(name, dim, "p p p", box, "meso", 1, create, cmds, nsteps, units="lj")


_add(_lj("lj3d", 3, 8))
_add(_lj("lj2d", 2, 80))
# the shipped shock-tube decks as they are: shrink-wrapped x (examples/USER/sph/shock_tube/shock{2d,3d}.lmp:2-3, boundary s p p)
_add(_shock("shock3d_shrink", 3, 45, bnd="s p p"))
_add(_shock("shock2d_shrink", 2, 40, bnd="ms p p", fill="-25.5"))
_add(_shock("gas3d_shrink", 3, 45, onetype=True, bnd="s p p"))


# ---- C3 scaled down: periodic two-phase box (square_to_sphere/droplet.lmp + cube.lmp) ----
def _droplet(name, dim, nx, nsteps, heat=None, skin=0.0, every=1, check="yes", static=False):
    L = 1.0; dx = L / nx; h = 3.0 * dx; rho = 1.0; c = 10.0; eta = 5e-2; alpha = 0.2; a = 0.2
    m = dx ** dim * rho
    if dim == 2:
        box = ((0, 0, -1e-3), (L, L, 1e-3)); lat = "sq"
        rsq = "region rsq block %s %s %s %s EDGE EDGE units box" % tuple(_f(v) for v in (0.5 - a, 0.5 + a, 0.5 - a, 0.5 + a))
    else:
        box = ((0, 0, 0), (L, L, L)); lat = "sc"
        rsq = "region rsq block %s %s %s %s %s %s units box" % tuple(_f(v) for v in (0.5 - a, 0.5 + a) * 3)
    create = """lattice %s %s
create_atoms 1 region box
%s
set region rsq type 2
set group all meso_rho %s
set group all mass %s
set type 1 meso_e 1.0
set type 2 meso_e 1.5
set type 1 meso_cv 1.0
set type 2 meso_cv 2.0
displace_atoms all random %s %s %s 4711 units box""" % (lat, _f(dx), rsq, _f(rho), _f(m), _f(0.05 * dx), _f(0.05 * dx),
                                                         _f(0.05 * dx) if dim == 3 else "0.0")
    subs = ["sph/rhosum/multiphase 1", "sph/colorgradient 1", "sph/taitwater/multiphase", "sph/surfacetension"]
    if heat:
        subs.append(heat)
    cmds = [("pair_style", "hybrid/overlay", *subs), ("pair_coeff", "* *", "sph/rhosum/multiphase", h),
            ("pair_coeff", "2 2", "sph/colorgradient", h, 0.0), ("pair_coeff", "1 2", "sph/colorgradient", h, alpha),
            ("pair_coeff", "1 1", "sph/colorgradient", h, 0.0),
            ("pair_coeff", "1 2", "sph/taitwater/multiphase", rho, c, eta, 7.0, h, 0.0),
            ("pair_coeff", "1 1", "sph/taitwater/multiphase", rho, c, eta, 7.0, h, 0.0),
            ("pair_coeff", "2 2", "sph/taitwater/multiphase", rho, c, eta, 7.0, h, 0.0),
            ("pair_coeff", "* *", "sph/surfacetension", h)]
    if heat == "sph/heatconduction/multiphase":
        cmds += [("pair_coeff", "* *", heat, 0.3, h)]
    elif heat == "sph/heatconduction/phasechange":
        cmds += [("pair_coeff", "1 1", heat, 0.2, h), ("pair_coeff", "1 2", heat, 0.3, h, "NULL", 1.2),
                 ("pair_coeff", "2 2", heat, 0.6, h)]
    dt = 0.0 if static else min(0.25 * dx / c, 0.125 * dx * dx / eta * rho)
    cmds += [("neighbor", skin), ("neigh_modify", dict(delay=0, every=every, check=check)), ("comm_modify", "yes"),
             ("timestep", dt), ("fix", "all", "meso")]
    return Case(name, dim, "p p p", box, "meso/multiphase", 2, create, cmds, nsteps)


# examples/USER/sph/poiseuille/poiseuille.lmp with its vars.lmp values (nx 5 -> 10, ny 30): reverse Poiseuille flow in a periodic 2-D box, one type,
# meso/multiphase, rhosum/multiphase + taitwater/multiphase, the body force as an atom-style variable that flips sign at Ly / 2, fix enforce2d
def _poiseuille(name, nsteps):
    nx, ny, Ly = 10, 30, 2e-3
    dx = Ly / ny; Lx = dx * nx; h = 3.0 * dx
    rho, c, eta, gx = 1e3, 1.25e-4, 1e-3, 1e-4
    m = dx ** 2 * rho
    dt = min(0.25 * h / c, 0.125 * h * h * rho / eta) * 0.25
    create = """lattice sq %s origin 0.5 0.5 0.0
create_atoms 1 region box
set group all meso_rho %s
set group all mass %s
displace_atoms all random %s %s 0.0 9731 units box""" % (_f(dx), _f(rho), _f(m), _f(0.02 * dx), _f(0.02 * dx))
    # (the shipped deck starts from the perfect lattice, where whole neighbor shells sit exactly on the cutoff and stay there -- rows move
    #  rigidly in x -- so that list membership hangs on the last bit of every coordinate; the 2 % jitter keeps the fixture well-posed)
    cmds = [("fix", "all", "meso"), ("neighbor", 0.0), ("neigh_modify", dict(delay=0, every=1)), ("comm_modify", "yes"),
            ("pair_style", "hybrid/overlay", "sph/rhosum/multiphase 1", "sph/taitwater/multiphase"),
            ("pair_coeff", "* *", "sph/taitwater/multiphase", rho, c, eta, 1.0, h, 0.0), ("pair_coeff", "* *", "sph/rhosum/multiphase", h),
            ("timestep", dt), ("variable", "bodyfx", "atom", "mass*%s*((y<%s/2.0)-(y>%s/2.0))" % (_f(gx), _f(Ly), _f(Ly))),
            ("fix", "all", "addforce", "v_bodyfx", 0.0, 0.0), ("fix", "all", "enforce2d")]
    return Case(name, 2, "p p p", ((0, 0, 0), (Lx, Ly, dx)), "meso/multiphase", 1, create, cmds, nsteps)


_add(_poiseuille("poiseuille2d", 60))
_add(_droplet("droplet2d", 2, 30, 40))
_add(_droplet("droplet3d", 3, 12, 20))
_add(_droplet("droplet3d_heat", 3, 12, 15, heat="sph/heatconduction/multiphase"))
_add(_droplet("droplet2d_pcheat_skin", 2, 30, 40, heat="sph/heatconduction/phasechange", skin=0.002, every=2))
# timestep 0: rho / colorgradient repeat every step, so the one-step-stale ghost values of the multiphase styles
#  equal the fresh ones and the result does not depend on the domain decomposition (multi-GPU check)
_add(_droplet("droplet3d_static", 3, 12, 3, heat="sph/heatconduction/phasechange", static=True))
_add(_droplet("droplet2d_static", 2, 30, 3, static=True))


# ---- C4 scaled down: random liquid box with a vapour seed, heat conduction + fix phase_change ----
def _bubble(name, dim, nx, nsteps, thermostat=False):
    L = 1.0; dx = L / nx; h = 3.0 * dx
    rho_l, rho_v = 1.0, 0.1
    c_v, c_l = 200.0 / math.sqrt(rho_v), 200.0 / math.sqrt(rho_l)
    eta_l, eta_v, alpha = 1.0, 0.69, 500.0
    D_l, D_v, cv_l, cv_v = 0.2, 0.6, 0.04, 0.06
    Hwv, Tc, Tinf = 8.0, 0.0, 1.0; Tt = Tc + 0.1
    m_v, m_l = dx ** dim * rho_v, dx ** dim * rho_l
    eta_ld = 2 * eta_l * eta_v / (eta_v + eta_l); D_ld = 2 * D_l * D_v / (D_v + D_l)
    box = ((0, 0, 0), (L, L, L if dim == 3 else dx))
    zc = 0.5 if dim == 3 else 0.0
    create = """lattice %s %s origin 0.5 0.5 %s
create_atoms 1 region box
displace_atoms all random %s %s %s 12345 units box
region rsq sphere 0.5 0.5 %s 0.22 units box
set region rsq type 2
set type 2 meso_cv %s
set type 1 meso_cv %s
set type 2 meso_e %s
set type 1 meso_e %s
set type 2 mass %s
set type 1 mass %s
set type 2 meso_rho %s
set type 1 meso_rho %s""" % ("sc" if dim == 3 else "sq", _f(dx), "0.5" if dim == 3 else "0", _f(0.2 * dx), _f(0.2 * dx),
                            _f(0.2 * dx) if dim == 3 else "0.0", _f(zc), _f(cv_v), _f(cv_l), _f(cv_v * 0.6), _f(cv_l * Tinf),
                            _f(m_v), _f(m_l), _f(rho_v), _f(rho_l))
    dts = [0.125 * dx * dx / eta_v * rho_v, 0.125 * dx * dx / eta_l * rho_l, 0.25 * math.sqrt(rho_v * dx ** 3 / (6.0 * alpha)),
           0.25 * dx / c_v, 0.25 * dx / c_l, 0.1 * 1.44 * rho_l * cv_l * dx * dx / D_l, 0.1 * 1.44 * rho_v * cv_v * dx * dx / D_v]
    hp = "sph/heatconduction/phasechange"
    cmds = [("pair_style", "hybrid/overlay", "sph/rhosum/multiphase 1", "sph/colorgradient 1", "sph/taitwater/multiphase",
             "sph/surfacetension", hp),
            ("pair_coeff", "* *", "sph/rhosum/multiphase", h),
            ("pair_coeff", "2 2", "sph/colorgradient", h, 0.0), ("pair_coeff", "1 2", "sph/colorgradient", h, alpha),
            ("pair_coeff", "1 1", "sph/colorgradient", h, 0.0),
            ("pair_coeff", "1 2", "sph/taitwater/multiphase", rho_l, c_l, eta_ld, 1.0, h, 0.0),
            ("pair_coeff", "1 1", "sph/taitwater/multiphase", rho_l, c_l, eta_l, 1.0, h, 0.0),
            ("pair_coeff", "2 2", "sph/taitwater/multiphase", rho_v, c_v, eta_v, 1.0, h, 0.0),
            ("pair_coeff", "* *", "sph/surfacetension", h),
            ("pair_coeff", "1 1", hp, D_l, h), ("pair_coeff", "1 2", hp, D_ld, h, "NULL", Tc), ("pair_coeff", "2 2", hp, D_v, h),
            ("neighbor", 0.0), ("neigh_modify", dict(delay=0, every=1)), ("comm_modify", "yes"), ("timestep", min(dts)),
            ("fix", "all", "meso"),
            ("fix", "bubble", "phase_change", Tc, Tt, Hwv, 0.5 * dx, m_v, h, 1, 2, 1, 123456, 0.05, "region", "box", "units", "box")]
    regions = ()
    if thermostat:      # bubble.lmp:97-103: far-field thermostat outside a sphere, fixed density in a corner block, 2-D constraint
        regions = (("rtemp", "sphere", 0.5, 0.5, zc, 0.5 - 2.6 * dx), ("rcorner", "block", "EDGE", 0.2, "EDGE", 0.2, "EDGE", "EDGE"))
        cmds += [("fix", "all", "setmeso", "meso_t", Tinf, "noregion", "rtemp"), ("fix", "all", "setmeso", "meso_rho", rho_l, "region", "rcorner")]
        if dim == 2:
            cmds += [("fix", "all", "enforce2d")]
    return Case(name, dim, "p p p", box, "meso/multiphase", 2, create, cmds, nsteps, groups=(("bubble", 2),), tol_traj=1e-8, regions=regions)


_add(_bubble("bubble2d", 2, 32, 30))
_add(_bubble("bubble3d", 3, 12, 12))
_add(_bubble("bubble2d_thermostat", 2, 32, 25, thermostat=True))
_add(_bubble("bubble3d_thermostat", 3, 12, 10, thermostat=True))


# ---- 1000-step trajectories with the DEFAULT atom_modify (spatial sort at every setup and on the first rebuild at or after step
#      1000, atom.cpp:63-65, verlet.cpp:106,251): the north star's "1e-6 after 1000 steps", and the test that the engine re-numbers
#      its local indices when Atom::sort does (output order, half-list ownership, fix phase_change draw order) ----
def _long(c, name, nsteps, tol):
    c.name, c.nsteps, c.tol_traj, c.sort = name, nsteps, tol, None
    return c


_add(_long(_dam("dam2d", 2, 60), "dam2d_1000", 1000, 1e-6))
_add(_long(_droplet("droplet2d", 2, 30, 40), "droplet2d_1000", 1000, 1e-6))
_add(_long(_bubble("bubble2d", 2, 32, 30), "bubble2d_1000", 1000, 1e-6))


# ---- lid-driven cavity (examples/USER/sph/cavity_flow/cavity_flow.lmp): sph/taitwater/morris ALONE in a periodic box, a strip of driver
#      particles that starts with a velocity and feels no force (fix setforce 0 0 0).  Two things no other deck has:
#      (1) no full-list sub-style, so the reference builds its half list with half_bin_newton (neigh_half_bin.cpp:285-420), whose pair
#          ownership is not half_from_full_newton's;
#      (2) non-zero velocities at setup: the ghosts carry the border-time copy of vest (zero) through the setup force evaluation, the owned
#          atoms the value FixMeso::setup_pre_force gave them (fix_meso.cpp:68-85), so every pair across a periodic face is evaluated with one
#          fresh and one stale velocity -- from the side of the atom that holds the pair.
#      `cavity2d_rhosum` is the same box behind hybrid/overlay sph/rhosum (full list -> half_from_full_newton): (2) without (1). ----
def _cavity(name, nsteps, rhosum=False):
    dx = 0.025e-3; h = 6.5e-5; L = 44 * dx
    box = ((0.0, 0.0, -1.0e-6), (L, L, 1.0e-6))
    m = 1000.0 * dx * dx
    create = """lattice sq %s
create_atoms 1 box
region strip block EDGE EDGE %s EDGE EDGE EDGE units box
set region strip type 2
set group all meso_rho 1000.0
mass 1 %s
mass 2 %s
group driver type 2
variable vx0 atom 0.0005*sin(6.283185307179586*y/%s)
variable vy0 atom 0.0003*cos(6.283185307179586*x/%s)
velocity all set v_vx0 v_vy0 0.0 units box
velocity driver set 0.001 0.0 0.0 units box""" % (_f(dx), _f(0.8 * L), _f(m), _f(m), _f(L), _f(L))
    if rhosum:
        ps = [("pair_style", "hybrid/overlay", "sph/rhosum 1", "sph/taitwater/morris"),
              ("pair_coeff", "* *", "sph/taitwater/morris", 1000.0, 0.1, 1.0e-3, h), ("pair_coeff", "* *", "sph/rhosum", h)]
    else:
        ps = [("pair_style", "sph/taitwater/morris"), ("pair_coeff", "* *", 1000.0, 0.1, 1.0e-3, h)]
    cmds = [("mass", "1", m), ("mass", "2", m)] + ps + [("neighbor", 3.0e-6), ("timestep", 5.0e-5),
            ("fix", "all", "meso"), ("fix", "driver", "setforce", 0.0, 0.0, 0.0)]
    return Case(name, 2, "p p p", box, "meso", 2, create, cmds, nsteps, groups=(("driver", 2),))


_add(_cavity("cavity2d", 60))
_add(_cavity("cavity2d_rhosum", 60, rhosum=True))


# cavity_flow.lmp as shipped, in miniature: three types (fluid, stationary walls, driver strip), `pair_style hybrid` with ONE sub-style and
# `pair_coeff 2 3 none` (walls and driver do not interact: their cutneighsq is 0, the sub-style's own setflag stays set -- the hybrid's map
# decides, pair_hybrid.cpp:378-398, 519-542), everything periodic so walls and driver meet their own images
def _cavity_none(name, nsteps):
    dx = 0.025e-3; h = 6.5e-5; L = 44 * dx
    box = ((0.0, 0.0, -1.0e-6), (L, L, 1.0e-6))
    m = 1000.0 * dx * dx
    create = """lattice sq %s
create_atoms 1 box
region rlow block EDGE EDGE EDGE %s EDGE EDGE units box
region rleft block EDGE %s EDGE EDGE EDGE EDGE units box
region rright block %s EDGE EDGE EDGE EDGE EDGE units box
region strip block EDGE EDGE %s EDGE EDGE EDGE units box
set region rlow type 2
set region rleft type 2
set region rright type 2
set region strip type 3
set group all meso_rho 1000.0
mass 1 %s
mass 2 %s
mass 3 %s
group driver type 3
velocity driver set 0.001 0.0 0.0 units box""" % (_f(dx), _f(3.2 * dx), _f(3.2 * dx), _f(L - 3.2 * dx), _f(L - 3.2 * dx), _f(m), _f(0.5 * m), _f(0.5 * m))
    cmds = [("mass", "1", m), ("mass", "2", 0.5 * m), ("mass", "3", 0.5 * m),
            ("pair_style", "hybrid", "sph/taitwater/morris"), ("pair_coeff", "* *", "sph/taitwater/morris", 1000.0, 0.1, 1.0e-3, h),
            ("pair_coeff", "2 3", "none"), ("neighbor", 3.0e-6), ("timestep", 5.0e-5),
            ("fix", "fluid", "meso"), ("fix", "driver", "meso"), ("fix", "walls", "meso/stationary"), ("fix", "driver", "setforce", 0.0, 0.0, 0.0)]
    return Case(name, 2, "p p p", box, "meso", 3, create, cmds, nsteps, groups=(("driver", 3), ("fluid", 1), ("walls", 2)))      # `driver` first: its group bit is taken when the create text defines it


_add(_cavity_none("cavity2d_none", 60))
# (The shipped deck also has TWINS: walls and driver are created separately, so where their regions overlap a wall atom and a driver atom
# sit on one lattice site.  `pair_coeff 2 3 none` gives the pair a zero neighbor cutoff but does not skip it in the sub-style's list
# (PairHybrid::init_style's mixing clause, pair_hybrid.cpp:459-462), so the twins (rsq = 0 <= 0) do interact -- with the sub-style's
# cutsq[2][3], which nothing ever initialises for an unassigned pair (PairHybrid::init_one only fills assigned ones; the array comes from
# malloc), and with cut[3][2] / viscosity[3][2] unset if the list holds the pair the other way round.  A fixture built that way gave NaN
# in the reference for one atom order and an interaction that lasts "until rsq exceeds the garbage" for the other, so there is none:
# the shipped deck itself is compared through the shells, where the shell reads the very arrays the reference would
# (tests/test_shell_shipped_cpu.py, tests/test_gpu_zz_shipped.py).)
