"""Host-side mirror of the reference's input-deck semantics for the SPH hot path.

Only what sits directly on the boundary of the path is mirrored: the commands
whose parsed state the `/b200` shells hand across the C-ABI.  Same command
names, argument order and error messages as the reference (file:line cited per
method, relative to /root/reference/), so the parity tests read like decks:

    d = Deck(dimension=3, boundary="p p p", box=((0,0,0),(1,1,1)), atom_style="meso/multiphase", ntypes=2)
    d.pair_style("hybrid/overlay", "sph/rhosum/multiphase 1", "sph/taitwater/multiphase")
    d.pair_coeff("* *", "sph/rhosum/multiphase", h)
    d.neighbor(0.0, "bin"); d.neigh_modify(delay=0, every=1)
    d.fix("integrate", "all", "meso")

It produces *tables* (Pair::init / PairHybrid::init_one / Neighbor::init output),
not physics: nothing here computes a force.
"""
import math
import numpy as np

PAIR_STYLES = {
    "sph/rhosum": 1, "sph/rhosum/multiphase": 2, "sph/taitwater": 3, "sph/taitwater/morris": 4,
    "sph/taitwater/multiphase": 5, "sph/colorgradient": 6, "sph/surfacetension": 7,
    "sph/heatconduction": 8, "sph/heatconduction/multiphase": 9, "sph/heatconduction/phasechange": 10, "sph/idealgas": 11,
    "sph/lj": 12,        # list-order dependent (pair_sph_lj.cpp:139): csrc/b200_lj.cuh
}
_NSETTINGS = {1: 1, 2: 1, 6: 1}           # styles whose settings() takes Nstep
_NCOEFF = {1: (1,), 2: (1,), 3: (4,), 4: (4,), 5: (6,), 6: (2,), 7: (1,), 8: (2,), 9: (2,), 10: (2, 4), 11: (2,), 12: (2,)}


class DeckError(RuntimeError):
    """error->all() of the reference (src/error.cpp:86-105)"""


def bounds(s, nmax, nmin=1):
    """Force::bounds, src/force.cpp:732-754"""
    s = str(s)
    if "*" not in s:
        lo = hi = int(s)
    elif len(s) == 1:
        lo, hi = nmin, nmax
    elif s[0] == "*":
        lo, hi = nmin, int(s[1:])
    elif s[-1] == "*":
        lo, hi = int(s[:-1]), nmax
    else:
        a, b = s.split("*")
        lo, hi = int(a), int(b)
    if lo < nmin or hi > nmax:
        raise DeckError("Numeric index is out of bounds")
    return lo, hi


class SubStyle:
    """One PairSPH* object: its coeff()/init_one() state (src/USER-SPH/pair_sph_*.cpp)."""

    def __init__(self, name, settings, ntypes):
        if name not in PAIR_STYLES:
            raise DeckError("Unknown pair style %s" % name)
        self.name, self.style, self.n = name, PAIR_STYLES[name], ntypes
        want = _NSETTINGS.get(self.style, 0)
        if len(settings) != want:
            raise DeckError("Illegal number of setting arguments for pair_style %s" % name)
        self.nstep = int(settings[0]) if want else 0
        n1 = ntypes + 1
        self.setflag = np.zeros((n1, n1), np.int32)
        self.cut = np.zeros((n1, n1)); self.cutsq = np.zeros((n1, n1))
        self.rho0 = np.zeros(n1); self.B = np.zeros(n1); self.soundspeed = np.zeros(n1)
        self.gamma = np.zeros(n1); self.rbackground = np.zeros(n1)
        self.viscosity = np.zeros((n1, n1)); self.alpha = np.zeros((n1, n1))
        self.tc = np.zeros((n1, n1)); self.fixflag = np.zeros((n1, n1), np.int32)
        self.mapped = np.zeros((n1, n1), np.int32)

    def coeff(self, I, J, args):
        """PairSPH*::coeff -- argument order per style:
        rhosum[/multiphase]  h                      pair_sph_rhosum.cpp:238-262
        taitwater[/morris]   rho0 c0 nu h           pair_sph_taitwater.cpp:238-276
        taitwater/multiphase rho0 c eta gamma h rb  pair_sph_taitwater_multiphase.cpp:226-263
        colorgradient        h alpha                pair_sph_colorgradient.cpp:216-252
        surfacetension       h                      pair_sph_surfacetension.cpp:225-250
        heatconduction[/mp]  D h                    pair_sph_heatconduction.cpp:168-195
        heatconduction/phasechange D h [Ti|NULL Tj|NULL]  ..._phasechange.cpp:177-225
        idealgas             nu h                   pair_sph_idealgas.cpp:210-238
        lj                   nu h                   pair_sph_lj.cpp:217-247"""
        if len(args) not in _NCOEFF[self.style]:
            raise DeckError("Incorrect args for pair_style %s coefficients" % self.name)
        ilo, ihi = bounds(I, self.n); jlo, jhi = bounds(J, self.n)
        st = self.style
        f = [None if (isinstance(a, str) and a == "NULL") else float(a) for a in args]
        ff_one, tc_one = 0, 0.0
        if st in (1, 2, 7):
            cut_one = f[0]
        elif st in (3, 4):
            rho0, c0, nu, cut_one = f
            B_one = c0 * c0 * rho0 / 7.0
        elif st == 5:
            rho0, c0, nu, gam, cut_one, rb = f
            B_one = c0 * c0 * rho0 / gam
        elif st in (11, 12):
            nu, cut_one = f
        elif st == 6:
            cut_one, alpha_one = f
        else:
            alpha_one, cut_one = f[0], f[1]
            if st == 10 and len(f) == 4:
                if f[2] is None and f[3] is not None:
                    ff_one, tc_one = 2, f[3]
                elif f[3] is None and f[2] is not None:
                    ff_one, tc_one = 1, f[2]
                else:
                    raise DeckError("Incorrect args for pair coefficients")
        count = 0
        for i in range(ilo, ihi + 1):
            if st in (3, 4, 5):
                self.rho0[i], self.soundspeed[i], self.B[i] = rho0, c0, B_one
                if st == 5:
                    self.gamma[i], self.rbackground[i] = gam, rb
            for j in range(max(jlo, i), jhi + 1):
                self.cut[i, j] = cut_one
                if st in (3, 4, 5, 11, 12):
                    self.viscosity[i, j] = nu
                if st in (6, 8, 9, 10):
                    self.alpha[i, j] = alpha_one
                if ff_one == 1:
                    self.fixflag[i, j], self.tc[i, j] = i, tc_one
                elif ff_one == 2:
                    self.fixflag[i, j], self.tc[i, j] = j, tc_one
                self.setflag[i, j] = 1
                count += 1
        if count == 0:
            raise DeckError("Incorrect args for pair coefficients")

    def init_one(self, i, j):
        """PairSPH*::init_one: symmetric copies, returns the cutoff"""
        if not self.setflag[i, j]:
            raise DeckError("All pair %s coeffs are not set" % self.name)
        self.cut[j, i] = self.cut[i, j]
        if self.style != 11:      # PairSPHIdealGas::init_one (pair_sph_idealgas.cpp:244-253) mirrors only cut: viscosity[j][i] stays
            self.viscosity[j, i] = self.viscosity[i, j]       # as allocated (zero) for j > i -- SURVEY Appendix B style quirk, reproduced
        self.alpha[j, i] = self.alpha[i, j]
        self.tc[j, i] = self.tc[i, j]
        self.fixflag[j, i] = self.fixflag[i, j]
        return self.cut[i, j]


class Deck:
    def __init__(self, dimension=3, boundary="p p p", box=((0, 0, 0), (1, 1, 1)), atom_style="meso",
                 ntypes=1, units="si", newton="on"):
        if atom_style not in ("meso", "meso/multiphase"):
            raise DeckError("Unknown atom style %s" % atom_style)
        if newton != "on":
            raise DeckError("the b200 SPH package requires newton on")
        self.dimension = int(dimension)
        b = boundary.split()
        if len(b) != 3:
            raise DeckError("Illegal boundary command")
        self.boundary = []        # Domain::set_boundary (domain.cpp:1440-1492): one letter = both faces, two = lo then hi
        for word in b:
            if len(word) not in (1, 2) or any(c not in "pfsm" for c in word):
                raise DeckError("Illegal boundary command")
            faces = [("pfsm").index(c) for c in (word if len(word) == 2 else word * 2)]
            if (faces[0] == 0) != (faces[1] == 0):
                raise DeckError("Both sides of boundary must be periodic")
            self.boundary.append(faces)
        self.periodicity = [1 if f[0] == 0 else 0 for f in self.boundary]
        self.shrink = any(v >= 2 for f in self.boundary for v in f)
        if self.dimension == 2 and not self.periodicity[2]:
            raise DeckError("Cannot use nonperiodic boundares with 2d simulation")  # sic, src/domain.cpp
        self.boxlo = [float(v) for v in box[0]]; self.boxhi = [float(v) for v in box[1]]
        # Domain::set_initial_box (domain.cpp:181-207): small from the box as given, s faces pushed out by it, m faces remember it
        self.small = [1.0e-4 * (hi - lo) for lo, hi in zip(self.boxlo, self.boxhi)]
        self.minbox = [[self.boxlo[d], self.boxhi[d]] for d in range(3)]
        for d in range(3):
            if self.boundary[d][0] == 2: self.boxlo[d] -= self.small[d]
            if self.boundary[d][1] == 2: self.boxhi[d] += self.small[d]
        self.multiphase = atom_style == "meso/multiphase"
        self.ntypes = int(ntypes)
        self.ftm2v = 1.0 if units in ("si", "lj", "cgs") else None
        if self.ftm2v is None:
            raise DeckError("b200 SPH package: units must be si, lj or cgs")
        self.mass_ = np.zeros(self.ntypes + 1)
        self.styles, self.hybrid = [], False
        self.skin, self.every, self.delay, self.check = 0.3, 1, 10, 1   # Neighbor::Neighbor defaults (neighbor.cpp:60-70)
        self.ghost_velocity = 0
        self.sortfreq, self.sort_binsize = 1000, 0.0      # atom_modify sort defaults (src/atom.cpp:63-65)
        self.dt, self.ntimestep = 0.005 if units == "lj" else 1.0e-8, 0   # Update::set_units (update.cpp)
        self.groups = {"all": 1}
        self.fixes = []
        self.variables = {}
        self.regions = {}
        self._initd = False

    # ---- simple commands ---------------------------------------------------
    def mass(self, I, value):
        lo, hi = bounds(I, self.ntypes)
        self.mass_[lo:hi + 1] = float(value)

    def timestep(self, dt):
        self.dt = float(dt)

    def neighbor(self, skin, style="bin"):
        if style != "bin":
            raise DeckError("b200 SPH package supports neighbor style bin only")
        self.skin = float(skin)

    def neigh_modify(self, every=None, delay=None, check=None):
        """Neighbor::modify_params (neighbor.cpp:2003-2045)"""
        if every is not None: self.every = int(every)
        if delay is not None: self.delay = int(delay)
        if check is not None: self.check = 1 if check in (1, True, "yes") else 0

    def atom_modify(self, sort=None):
        """atom_modify sort Nfreq binsize (Atom::modify_params, src/atom.cpp:540-552)"""
        if sort is not None:
            freq, binsize = int(sort[0]), float(sort[1])
            if freq < 0 or binsize < 0.0:
                raise DeckError("Illegal atom_modify command")
            self.sortfreq, self.sort_binsize = freq, binsize

    def comm_modify(self, vel="no"):
        self.ghost_velocity = 1 if vel in (1, True, "yes") else 0

    def group(self, name):
        """returns the group bit (Group::assign allocates the next free bit)"""
        if name not in self.groups:
            if len(self.groups) >= 32:
                raise DeckError("Too many groups")
            self.groups[name] = 1 << len(self.groups)
        return self.groups[name]

    def region(self, ID, style, *args):
        """region ID block xlo xhi ylo yhi zlo zhi | sphere xc yc zc r  (units box, side in; EDGE = box bound)"""
        if style == "block":
            v = []
            for k, a in enumerate(args[:6]):
                if a == "EDGE":
                    v.append(-1.0e20 if k % 2 == 0 else 1.0e20)      # region_block.cpp: EDGE -> -BIG / BIG
                else:
                    v.append(float(a))
            self.regions[ID] = (1, v)
        elif style == "sphere":
            self.regions[ID] = (2, [float(a) for a in args[:4]] + [0.0, 0.0])
        else:
            raise DeckError("b200 SPH package: region styles block and sphere are supported")

    # ---- pair styles -------------------------------------------------------
    def pair_style(self, name, *sub):
        """pair_style <style> args | pair_style hybrid/overlay <sub1 args> <sub2 args> ...
        (PairHybrid::settings, pair_hybrid.cpp:190-258)"""
        self.styles, self._initd = [], False
        n1 = self.ntypes + 1
        self.hset = np.zeros((n1, n1), np.int32)                 # PairHybrid::setflag
        self.hmap = [[[] for _ in range(n1)] for _ in range(n1)]   # PairHybrid::nmap / map (pair_hybrid.h:59-60)
        self.overlay = name == "hybrid/overlay"
        if name in ("hybrid/overlay", "hybrid"):
            self.hybrid = True
            for s in sub:
                w = s.split()
                self.styles.append(SubStyle(w[0], w[1:], self.ntypes))
            names = [s.name for s in self.styles]
            if len(set(names)) != len(names):
                raise DeckError("b200 SPH package: repeated hybrid sub-styles are not supported")
        else:
            self.hybrid = False
            self.styles.append(SubStyle(name, [str(a) for a in sub], self.ntypes))

    def pair_coeff(self, IJ, *args):
        """pair_coeff I J [substyle] args  (PairHybrid::coeff, pair_hybrid.cpp:330-401)"""
        I, J = IJ.split()
        self._initd = False
        if self.hybrid:
            sub, args = args[0], args[1:]
            m = [k for k, s in enumerate(self.styles) if s.name == sub]
            none = not m and sub == "none"
            if not m and not none:
                raise DeckError("Pair coeff for hybrid has invalid style")
            if not none:
                self.styles[m[0]].coeff(I, J, args)
            # which type pairs map to which sub-style: `none` wipes the map, plain hybrid replaces it (pair_hybrid.cpp:378-398),
            # hybrid/overlay adds the sub-style if it is new for the pair (pair_hybrid_overlay.cpp:88-104)
            ilo, ihi = bounds(I, self.ntypes); jlo, jhi = bounds(J, self.ntypes)
            count = 0
            for i in range(ilo, ihi + 1):
                for j in range(max(jlo, i), jhi + 1):
                    if none:
                        self.hset[i, j] = 1; self.hmap[i][j] = []; count += 1
                    elif self.styles[m[0]].setflag[i, j]:
                        if self.overlay:
                            if m[0] not in self.hmap[i][j]:
                                self.hmap[i][j].append(m[0])
                        else:
                            self.hmap[i][j] = [m[0]]
                        self.hset[i, j] = 1; count += 1
            if count == 0:
                raise DeckError("Incorrect args for pair coefficients")
        else:
            self.styles[0].coeff(I, J, args)

    # ---- variables ---------------------------------------------------------
    def variable(self, name, style, formula):
        """variable name equal|atom formula (variable.cpp:136-330): kept as text; `v_other` references are spliced in
        (in parentheses) when a fix takes the variable, ${} is the caller's business as it is the input parser's in the reference"""
        if style not in ("equal", "atom"):
            raise DeckError("b200 SPH package: variable styles equal and atom")
        self.variables[name] = str(formula)

    def _formula(self, vname, depth=0):
        import re
        name = str(vname)[2:]
        if name not in self.variables:
            raise DeckError("Variable name for fix does not exist")
        if depth > 8:
            raise DeckError("Variable has circular dependency")
        return re.sub(r"\bv_(\w+)", lambda m: "(" + self._formula(m.group(0), depth + 1) + ")", self.variables[name])

    # ---- fixes -------------------------------------------------------------
    def fix(self, ID, group, style, *args):
        """fix ID group style args: meso (fix_meso.cpp:38-49), meso/stationary, gravity
        (`gravity <mag> vector x y z`, fix_gravity.cpp:40-130,304-337), phase_change
        (fix_phase_change.cpp:46-125)."""
        if group not in self.groups:
            raise DeckError("Could not find fix group ID")
        bit = self.groups[group]
        if style in ("meso", "meso/stationary"):
            if args:
                raise DeckError("Illegal number of arguments for fix %s command" % style)
            self.fixes.append((style, bit, None))
        elif style == "gravity":
            mag = float(args[0])
            if args[1] != "vector":
                raise DeckError("b200 SPH package: fix gravity supports the vector style")
            xd, yd, zd = (float(a) for a in args[2:5])
            if self.dimension == 3:
                ln = math.sqrt(xd * xd + yd * yd + zd * zd); g = (xd / ln, yd / ln, zd / ln)
            else:
                ln = math.sqrt(xd * xd + yd * yd); g = (xd / ln, yd / ln, 0.0)
            self.fixes.append((style, bit, tuple(mag * c for c in g)))
        elif style == "enforce2d":
            if self.dimension == 3:
                raise DeckError("Cannot use fix enforce2d with 3d simulation")
            self.fixes.append((style, bit, None))
        elif style == "dt/reset":      # fix_dt_reset.cpp:40-98: N Tmin Tmax Xmax [units box]
            if len(args) < 4:
                raise DeckError("Illegal fix dt/reset command")
            nevery = int(args[0])
            minb, tmin = (0, 0.0) if str(args[1]) == "NULL" else (1, float(args[1]))
            maxb, tmax = (0, 0.0) if str(args[2]) == "NULL" else (1, float(args[2]))
            xmax = float(args[3])
            if list(args[4:]) != ["units", "box"]:
                raise DeckError("b200 SPH package: fix dt/reset needs `units box`")
            if nevery <= 0 or xmax <= 0.0 or (minb and tmin < 0.0) or (maxb and tmax < 0.0) or (minb and maxb and tmin >= tmax):
                raise DeckError("Illegal fix dt/reset command")
            self.fixes.append((style, bit, (nevery, minb, tmin, maxb, tmax, xmax)))
        elif style == "setmesode":     # fix_setmesode.cpp:38-78: value [region ID], constant value
            if not args or str(args[0]).startswith("v_") or str(args[0]) == "NULL":
                raise DeckError("b200 SPH package: fix setmesode supports a constant value")
            kind, reg = 0, [0.0] * 6
            if len(args) > 1:
                if args[1] != "region" or len(args) < 3:
                    raise DeckError("Illegal fix setmesode command")
                if args[2] not in self.regions:
                    raise DeckError("Region ID for fix setmesode does not exist")
                kind, reg = self.regions[args[2]]
            self.fixes.append((style, bit, (float(args[0]), kind, list(reg))))
        elif style == "setforce":      # fix_setforce.cpp:40-110, constant values or NULL
            if len(args) != 3:
                raise DeckError("Illegal fix setforce command")
            if any(str(a).startswith("v_") for a in args):
                raise DeckError("b200 SPH package: fix setforce supports constant values")
            sets = [0 if str(a) == "NULL" else 1 for a in args]
            vals = [0.0 if str(a) == "NULL" else float(a) for a in args]
            self.fixes.append((style, bit, (sets, vals)))
        elif style == "setmeso":
            which = {"meso_rho": 0, "meso_e": 1, "meso_t": 2}.get(args[0])
            if which is None:
                raise DeckError("Illegal fix setmeso command, meso_rho or meso_e must be given")
            formula = self._formula(args[1]) if str(args[1]).startswith("v_") else None
            kind, reg, inside = 0, [0.0] * 6, 1
            if len(args) > 2:
                if args[2] not in ("region", "noregion") or len(args) < 4:
                    raise DeckError("Illegal fix setmesode command")
                if args[3] not in self.regions:
                    raise DeckError("Region ID for fix setmesode does not exist")
                kind, reg = self.regions[args[3]]
                inside = 1 if args[2] == "region" else 0
            if formula is not None:    # varflag EQUAL / ATOM (fix_setmeso.cpp:100-140): the variable's formula goes to the device
                self.fixes.append(("setmeso/var", bit, (which, formula, kind, list(reg), inside)))
            else:
                self.fixes.append((style, bit, (which, float(args[1]), kind, list(reg), inside)))
        elif style == "addforce":      # fix_addforce.cpp:40-110: fx fy fz, each a constant or v_name (this mirror takes no keywords; the C++ shell composes every / region into the formula)
            if len(args) != 3:
                raise DeckError("b200 SPH package: fix addforce supports `fx fy fz` without keywords")
            vals = [0.0 if str(a).startswith("v_") else float(a) for a in args]
            forms = [self._formula(a) if str(a).startswith("v_") else None for a in args]
            self.fixes.append((style, bit, (vals, forms)))
        elif style == "phase_change":
            a = list(args)
            if len(a) < 11:
                raise DeckError("Illegal fix phase_change command")
            pc = dict(Tc=float(a[0]), Tt=float(a[1]), Hwv=float(a[2]), dr=float(a[3]), to_mass=float(a[4]),
                      cutoff=float(a[5]), from_type=int(a[6]), to_type=int(a[7]), nfreq=int(a[8]), seed=int(a[9]),
                      energy_chance_flag=0, change_chance=0.0, phase_change_rate=0.0, maxattempt=10)
            if pc["seed"] <= 0:
                raise DeckError("Illegal value for seed")
            m = 10
            if a[m] == "ENERGY":
                pc["energy_chance_flag"], pc["phase_change_rate"] = 1, float(a[m + 1]); m += 2
            else:
                pc["change_chance"] = float(a[m]); m += 1
                if pc["change_chance"] < 0:
                    raise DeckError("Illegal value for change_chance")
            region = None
            while m < len(a):
                if a[m] == "region": region = a[m + 1]; m += 2
                elif a[m] == "attempt": pc["maxattempt"] = int(a[m + 1]); m += 2
                elif a[m] == "units":
                    if a[m + 1] != "box": raise DeckError("Illegal fix phase_change command")
                    m += 2
                else:
                    raise DeckError("Illegal fix phase_change command")
            if region is None:
                raise DeckError("Must specify a region in fix phase_change")
            pc["first_step"] = self.ntimestep + 1       # next_reneighbor = update->ntimestep + 1 (:120)
            self.fixes.append((style, bit, pc))
        else:
            raise DeckError("Unknown fix style %s (b200 SPH package)" % style)

    # ---- init --------------------------------------------------------------
    def init(self):
        """Pair::init (pair.cpp:174-235) + PairHybrid::init_style/init_one
        (pair_hybrid.cpp:407-543) + Neighbor::init cutoffs (neighbor.cpp:236-282)
        + CommBrick::setup cutghost (comm_brick.cpp:166-172)."""
        n = self.ntypes
        n1 = n + 1
        nmap = np.zeros((n1, n1), np.int32)
        for s in self.styles:
            s.mapped[:] = 0
        if self.hybrid:
            used = {k for i in range(1, n1) for j in range(i, n1) for k in self.hmap[i][j]}
            for k, s in enumerate(self.styles):      # pair_hybrid.cpp:425-433: every sub-style must be mapped somewhere
                if k not in used:
                    raise DeckError("Pair hybrid sub-style is not used")
            for i in range(1, n1):
                if not self.hset[i, i]:
                    raise DeckError("All pair coeffs are not set")
        else:
            for i in range(1, n1):
                if not self.styles[0].setflag[i, i]:
                    raise DeckError("All pair coeffs are not set")
        self.cutsq = np.zeros((n1, n1))
        for i in range(1, n1):
            for j in range(i, n1):
                if self.hybrid:
                    if not self.hset[i, j]:
                        # mixing needs I,I and J,J on one identical single sub-style, which then
                        # fails in PairSPH*::init_one (no mixing rule) -> same message either way
                        raise DeckError("All pair coeffs are not set")
                    mapped = [self.styles[k] for k in self.hmap[i][j]]      # empty after `pair_coeff I J none`: cutoff 0 (pair_hybrid.cpp:519-542)
                    # ... but NOT skipped in the list of a sub-style that owns both I,I and J,J alone (PairHybrid::init_style's "mixing will
                    # assign this pair" clause, pair_hybrid.cpp:459-462): atoms of such types at the same position (rsq = 0 <= 0) interact
                    if not mapped and len(self.hmap[i][i]) == 1 and self.hmap[i][i] == self.hmap[j][j]:
                        s = self.styles[self.hmap[i][i][0]]
                        if s.setflag[i, j]:
                            cut = s.init_one(i, j)
                            s.cutsq[i, j] = s.cutsq[j, i] = cut * cut
                        s.mapped[i, j] = s.mapped[j, i] = 1
                else:
                    mapped = [s for s in self.styles if s.setflag[i, j]]
                    if not mapped:
                        raise DeckError("All pair coeffs are not set")
                cutmax = 0.0
                for s in mapped:
                    cut = s.init_one(i, j)
                    s.cutsq[i, j] = s.cutsq[j, i] = cut * cut
                    s.mapped[i, j] = s.mapped[j, i] = 1
                    cutmax = max(cutmax, cut)
                nmap[i, j] = nmap[j, i] = len(mapped)
                self.cutsq[i, j] = self.cutsq[j, i] = cutmax * cutmax
        self.cutneighsq = np.zeros((n1, n1))
        cutneighmax = 0.0
        for i in range(1, n1):
            for j in range(1, n1):
                cutoff = math.sqrt(self.cutsq[i, j])
                cut = cutoff + (self.skin if cutoff > 0.0 else 0.0)
                self.cutneighsq[i, j] = cut * cut
                cutneighmax = max(cutneighmax, cut)
        self.cutneighmax = cutneighmax
        self.cutghost = cutneighmax          # MAX(cutneighmax, cutghostuser = 0)
        self._initd = True
        return self
