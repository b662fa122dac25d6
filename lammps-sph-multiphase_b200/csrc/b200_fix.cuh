// b200_fix.cuh -- per-atom stages: fix meso / meso/stationary / gravity, reverse
// accumulation of ghost contributions, and the rebuild trigger.
#pragma once
#include "b200_expr.cuh"
#include "b200_common.cuh"

struct StepArrays {
  double4 *xt, *vr, *vm, *fd;
  double *e, *de;
  const int *mask;
  const double *cv;
  const int *tag;
};

// FixMeso::setup_pre_force (fix_meso.cpp:68-85): vest = v for atoms of the fix group
__global__ void k_setup_pre_force(int nlocal, FixList fl, StepArrays a)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  int m = a.mask[i];
  for (int k = 0; k < fl.n; k++)
    if (fl.kind[k] == 1 && (m & fl.bit[k])) {
      double4 v = a.vm[i], vr = a.vr[i];
      vr.x = v.x; vr.y = v.y; vr.z = v.z;
      a.vr[i] = vr;
    }
}

// Is FixMeso::setup_pre_force about to change vest for any owned atom?  Then the ghosts this setup's borders made carry a vest their
// owners no longer have through the setup force evaluation (Verlet::setup, verlet.cpp:106-127: borders, then setup_pre_force, then
// the pair styles; the next forward_comm comes with step 1) -- non-zero initial velocities, or a run that continues an earlier one.
__global__ void k_vest_stale(int nlocal, FixList fl, StepArrays a, int *flag)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  int m = a.mask[i];
  for (int k = 0; k < fl.n; k++)
    if (fl.kind[k] == 1 && (m & fl.bit[k])) {
      double4 v = a.vm[i], vr = a.vr[i];
      if (v.x != vr.x || v.y != vr.y || v.z != vr.z) *flag = 1;
      return;
    }
}

// modify->initial_integrate: FixMeso::initial_integrate (fix_meso.cpp:91-140) and
// largest squared displacement since the build of the atoms (owned or ghost) a cell held at the build, as fp32 bits rounded up
// (non-negative floats order like their bit patterns): the tile path decides per tile whether its mid / far rows are due
__device__ __forceinline__ void cell_disp_max(unsigned *celld, int cell, double dsq)
{
  const unsigned b = __float_as_uint(__double2float_ru(dsq));
  if (b > *(volatile unsigned *)(celld + cell)) atomicMax(celld + cell, b);
}
// FixMesoStationary::initial_integrate (fix_meso_stationary.cpp:71-92), fixes in deck order;
// plus Neighbor::check_distance's per-atom test (neighbor.cpp:1396-1404) on the new positions.
// dtp != NULL: the timestep lives on the device (fix dt/reset); dtp[0] = dt, dtf_per_dt = 0.5 ftm2v
__global__ void k_initial_integrate(int nlocal, FixList fl, StepArrays a, double dtv, double dtf, int check,
                                    const double *xhold, double triggersq, int *flag, int track, unsigned long long *dmaxsq, const double *dtp, double dtf_per_dt,
                                    const int *rowcell, unsigned *celld)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (dtp) { dtv = *dtp; dtf = dtf_per_dt * dtv; }      // dtf = 0.5 dt ftm2v
  int m = a.mask[i];
  double4 x = a.xt[i], vr = a.vr[i], v = a.vm[i], f = a.fd[i];
  double e = a.e[i], de = a.de[i];
  bool moved = false, touched = false;
  for (int k = 0; k < fl.n; k++) {
    if (fl.kind[k] > 2 || !(m & fl.bit[k])) continue;
    touched = true;
    e += dtf * de;            // half-step update of particle internal energy
    vr.w += dtf * f.w;        // ... and density
    if (fl.kind[k] == 1) {
      double dtfm = dtf / v.w;
      vr.x = v.x + 2.0 * dtfm * f.x; vr.y = v.y + 2.0 * dtfm * f.y; vr.z = v.z + 2.0 * dtfm * f.z;
      v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
      x.x += dtv * v.x; x.y += dtv * v.y; x.z += dtv * v.z;
      moved = true;
    }
  }
  if (touched) { a.e[i] = e; a.vr[i] = vr; }
  if (moved) { a.vm[i] = v; a.xt[i] = x; }
  if (check || track) {
    double dx = x.x - xhold[3 * i], dy = x.y - xhold[3 * i + 1], dz = x.z - xhold[3 * i + 2];
    double dsq = dx * dx + dy * dy + dz * dz;
    if (check && dsq > triggersq) *flag = 1;
    if (track && moved) {      // largest displacement since the build (non-negative doubles order like their bit patterns)
      unsigned long long b = (unsigned long long)__double_as_longlong(dsq);
      if (b > *(volatile unsigned long long *)dmaxsq) atomicMax(dmaxsq, b);
      if (celld) cell_disp_max(celld, rowcell[i], dsq);
    }
  }
}
// far rows must be scanned once 2*dmax >= margin  (b200_neigh.cuh k_build)
// (tile rows: a mid zone between cut + margin/4 and cut + margin is scanned once 2*dmax >= margin/4; flag at scan_far[2])
__global__ void k_far_flag(const unsigned long long *dmaxsq, double marginsq, double midmarginsq, int *scan_far)
{
  const double d = 4.0 * __longlong_as_double((long long)*dmaxsq);
  scan_far[0] = d >= 0.99 * marginsq; scan_far[2] = d >= 0.99 * midmarginsq;
}

// (comm->reverse_comm runs before this kernel: b200_comm.cuh) modify->post_force (FixGravity::post_force, fix_gravity.cpp:262-295), then modify->final_integrate
// (fix_meso.cpp:144-180, fix_meso_stationary.cpp:96-112).  The three stages can be run fused (one
// pass over the owned atoms) or one by one for the stage-level ABI.
// progs / step / dt: the compiled variable formulas of fix addforce / setmeso (b200_expr.cuh) and the thermo keywords they may name
__global__ void k_post_final(int nlocal, FixList fl, StepArrays a, double dtf, int do_post, int do_final, const double *dtp, double dtf_per_dt,
                             const ExprProg *progs, double step, double dt)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (dtp) { dtf = dtf_per_dt * *dtp; dt = *dtp; }
  // what a formula can name about atom i (Variable::eval_tree: x, v, f as they stand when the fix runs; mass = rmass or mass[type])
  auto expr_in = [&](const double4 &fnow, const double4 &vnow) {
    const double4 x = a.xt[i];
    ExprIn in;
    in.x = x.x; in.y = x.y; in.z = x.z; in.vx = vnow.x; in.vy = vnow.y; in.vz = vnow.z; in.fx = fnow.x; in.fy = fnow.y; in.fz = fnow.z;
    in.mass = vnow.w; in.type = tw_type(__double_as_longlong(x.w)); in.id = a.tag[i]; in.step = step; in.dt = dt; in.time = 0.0;
    return in;
  };
  double4 f = a.fd[i];
  double de = a.de[i];
  bool fdirty = false;
  int m = a.mask[i];
  double4 v = a.vm[i];
  bool vdirty = false;
  if (do_post)
    for (int k = 0; k < fl.n; k++) {
      if (!(m & fl.bit[k])) continue;
      if (fl.kind[k] == 3) {                                  // FixGravity::post_force
        f.x += v.w * fl.acc[k][0]; f.y += v.w * fl.acc[k][1]; f.z += v.w * fl.acc[k][2];
        fdirty = true;
      } else if (fl.kind[k] == 4) {                           // FixSetMeso::post_force, constant value (fix_setmeso.cpp:211-236)
        int rk = fl.ipar[k][1];
        if (rk) {
          double4 x = a.xt[i];
          const double *r = &fl.par[k][1];
          bool in;
          if (rk == 1) in = x.x >= r[0] && x.x <= r[1] && x.y >= r[2] && x.y <= r[3] && x.z >= r[4] && x.z <= r[5];
          else { double dx = x.x - r[0], dy = x.y - r[1], dz = x.z - r[2]; in = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) <= r[3]; }
          if (in != (fl.ipar[k][2] != 0)) continue;
        }
        double val = fl.par[k][0];
        if (fl.prog[k][0]) { const ExprIn in = expr_in(f, v); val = expr_eval(progs[fl.prog[k][0] - 1], in); }      // varflag ATOM / EQUAL (fix_setmeso.cpp:238-262)
        if (fl.ipar[k][0] == 0) { double4 vr = a.vr[i]; vr.w = val; a.vr[i] = vr; }
        else a.e[i] = fl.ipar[k][0] == 2 ? a.cv[i] * val : val;
      } else if (fl.kind[k] == 5) {                           // FixEnforce2D::post_force (fix_enforce2d.cpp:77-89)
        v.z = 0.0; f.z = 0.0; fdirty = true; vdirty = true;
      } else if (fl.kind[k] == 6) {                           // FixSetForce::post_force, constant values (fix_setforce.cpp:241-251)
        if (fl.ipar[k][0]) f.x = fl.par[k][0];
        if (fl.ipar[k][1]) f.y = fl.par[k][1];
        if (fl.ipar[k][2]) f.z = fl.par[k][2];
        fdirty = true;
      } else if (fl.kind[k] == 8) {                           // FixAddForce::post_force (fix_addforce.cpp:269-320): all three components are evaluated before any is added
        double ax = fl.acc[k][0], ay = fl.acc[k][1], az = fl.acc[k][2];
        if (fl.prog[k][0] | fl.prog[k][1] | fl.prog[k][2]) {
          const ExprIn in = expr_in(f, v);
          if (fl.prog[k][0]) ax = expr_eval(progs[fl.prog[k][0] - 1], in);
          if (fl.prog[k][1]) ay = expr_eval(progs[fl.prog[k][1] - 1], in);
          if (fl.prog[k][2]) az = expr_eval(progs[fl.prog[k][2] - 1], in);
        }
        f.x += ax; f.y += ay; f.z += az;
        fdirty = true;
      }
    }
  if (fdirty) a.fd[i] = f;
  if (vdirty) a.vm[i] = v;
  if (do_post)
    for (int k = 0; k < fl.n; k++)
      if (fl.kind[k] == 7 && (m & fl.bit[k])) {             // FixSetMesodE::post_force, constant value (fix_setmesode.cpp:171-199)
        int rk = fl.ipar[k][1];
        if (rk) {
          double4 x = a.xt[i];
          const double *r = &fl.par[k][1];
          bool in;
          if (rk == 1) in = x.x >= r[0] && x.x <= r[1] && x.y >= r[2] && x.y <= r[3] && x.z >= r[4] && x.z <= r[5];
          else { double dx = x.x - r[0], dy = x.y - r[1], dz = x.z - r[2]; in = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))) <= r[3]; }
          if (!in) continue;
        }
        de = fl.par[k][0]; a.de[i] = de;
      }
  if (do_final) {
    double4 vr = a.vr[i];
    double e = a.e[i];
    bool vd = false, ed = false;
    for (int k = 0; k < fl.n; k++) {
      if (fl.kind[k] > 2 || !(m & fl.bit[k])) continue;
      if (fl.kind[k] == 1) {
        double dtfm = dtf / v.w;
        v.x += dtfm * f.x; v.y += dtfm * f.y; v.z += dtfm * f.z;
        vd = true;
      }
      e += dtf * de;
      vr.w += dtf * f.w;
      ed = true;
    }
    if (vd) a.vm[i] = v;
    if (ed) { a.e[i] = e; a.vr[i] = vr; }
  }
}

// FixDtReset::end_of_step (fix_dt_reset.cpp:131-186): per-atom largest timestep that keeps the displacement below xmax; the minimum over
// the group is taken on the bit patterns (positive doubles order like uint64)
__global__ void k_dt_min(int nlocal, int groupbit, const int *mask, const double4 *vm, const double4 *fd, double xmax, double ftm2v,
                         unsigned long long *dtmin_bits)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal || !(mask[i] & groupbit)) return;
  const double4 v = vm[i], f = fd[i];
  const double massinv = 1.0 / v.w;
  const double vsq = __dadd_rn(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)), __dmul_rn(v.z, v.z));
  const double fsq = __dadd_rn(__dadd_rn(__dmul_rn(f.x, f.x), __dmul_rn(f.y, f.y)), __dmul_rn(f.z, f.z));
  double dtv = 1.0e20, dtf = 1.0e20;
  if (vsq > 0.0) dtv = xmax / sqrt(vsq);
  if (fsq > 0.0) dtf = sqrt(2.0 * xmax / (ftm2v * sqrt(fsq) * massinv));
  double dt = fmin(dtv, dtf);
  const double dtsq = dt * dt;
  const double delx = __dadd_rn(__dmul_rn(dt, v.x), __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.5, dtsq), massinv), f.x), ftm2v));
  const double dely = __dadd_rn(__dmul_rn(dt, v.y), __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.5, dtsq), massinv), f.y), ftm2v));
  const double delz = __dadd_rn(__dmul_rn(dt, v.z), __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(0.5, dtsq), massinv), f.z), ftm2v));
  const double delr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(delx, delx), __dmul_rn(dely, dely)), __dmul_rn(delz, delz)));
  if (delr > xmax) dt *= xmax / delr;
  const unsigned long long b = (unsigned long long)__double_as_longlong(dt);
  if (b < *(volatile unsigned long long *)dtmin_bits) atomicMin(dtmin_bits, b);
}
// dtp[0] = dt, dtp[2] = Update::atime, ((long long *)dtp)[3] = Update::atimestep, ((long long *)dtp)[4] = FixDtReset::laststep
// (fix_dt_reset.cpp:175-181: nothing moves when the timestep did not change; else update_time() with the OLD dt, update.cpp:480-484)
__global__ void k_dt_apply(const unsigned long long *dtmin_bits, int minbound, double tmin, int maxbound, double tmax, double *dtp, long long step)
{
  double dt = __longlong_as_double((long long)*dtmin_bits);
  if (minbound) dt = fmax(dt, tmin);
  if (maxbound) dt = fmin(dt, tmax);
  const double old = dtp[0];
  if (dt == old) return;
  long long *lp = (long long *)dtp;
  dtp[2] = __dadd_rn(dtp[2], __dmul_rn((double)(step - lp[3]), old));
  lp[3] = step;
  lp[4] = step;
  dtp[0] = dt;
}

// Pair::virial_fdotr_compute (pair.cpp:1403-1451): sum over owned + ghost atoms of x (x) f (xx yy zz xy xz yz), taken before
// the reverse halo.  Deterministic two-stage reduction: fixed thread -> element assignment, fixed tree, then the block partials in order.
// v6 != NULL: sum the rows of v6[n][6] instead (per-row pair sums of the single-phase tile path).
#define VIR_BLOCKS 256
#define VIR_THREADS 256
__global__ void __launch_bounds__(VIR_THREADS) k_virial_partial(int n, const double4 *xt, const double4 *fd, const double *v6, double *partial)
{
  __shared__ double sh[6][VIR_THREADS];
  double w[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * VIR_THREADS + threadIdx.x; i < n; i += VIR_BLOCKS * VIR_THREADS) {
    if (v6) { for (int k = 0; k < 6; k++) w[k] += v6[(size_t)i * 6 + k]; }
    else {
      const double4 x = xt[i], f = fd[i];
      w[0] += x.x * f.x; w[1] += x.y * f.y; w[2] += x.z * f.z; w[3] += x.x * f.y; w[4] += x.x * f.z; w[5] += x.y * f.z;
    }
  }
  for (int k = 0; k < 6; k++) sh[k][threadIdx.x] = w[k];
  __syncthreads();
  for (int o = VIR_THREADS / 2; o; o >>= 1) {
    if (threadIdx.x < o) for (int k = 0; k < 6; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 6) partial[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void k_virial_final(const double *partial, double *out)
{
  int k = threadIdx.x;
  if (k >= 6) return;
  double s = 0.0;
  for (int b = 0; b < VIR_BLOCKS; b++) s += partial[b * 6 + k];
  out[k] = s;
}

// ---- host <-> device layout conversion (LAMMPS AoS per-atom arrays <-> packed double4 records) ----
struct HostMirror {            // device staging copies of the caller's arrays (NULL = field absent)
  double *x, *v, *vest, *f, *cg, *rho, *drho, *e, *de, *cv, *rmass;
  int *type, *mask, *tag;
};
struct PackArrays {
  double4 *xt, *vr, *vm, *fd, *cgm;
  double *e, *de, *cv;
  int *tag, *mask, *orig;
};
// bad[0]: a type is out of range; bad[1]: largest tag (warp maximum first, one atomic per warp)
__global__ void k_pack_atoms(int n, HostMirror m, PackArrays a, int multiphase, const double *mass, int ntypes, int *bad)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int tg = (i < n) ? (m.tag ? m.tag[i] : i + 1) : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) tg = max(tg, __shfl_xor_sync(FULLMASK, tg, o));
  if ((threadIdx.x & 31) == 0 && tg > *(volatile int *)(bad + 1)) atomicMax(bad + 1, tg);
  if (i >= n) return;
  int t = m.type[i];
  if (t < 1 || t > ntypes) { bad[0] = 1; t = 1; }
  double ms = (multiphase && m.rmass) ? m.rmass[i] : mass[t];
  const double *v = m.v ? m.v + 3 * i : nullptr, *ve = m.vest ? m.vest + 3 * i : v;
  a.xt[i] = make_double4(m.x[3 * i], m.x[3 * i + 1], m.x[3 * i + 2], __longlong_as_double((long long)pack_tw(t, 0, 0, 0)));
  a.vm[i] = make_double4(v ? v[0] : 0.0, v ? v[1] : 0.0, v ? v[2] : 0.0, ms);
  a.vr[i] = make_double4(ve ? ve[0] : 0.0, ve ? ve[1] : 0.0, ve ? ve[2] : 0.0, m.rho ? m.rho[i] : 0.0);
  a.fd[i] = make_double4(m.f ? m.f[3 * i] : 0.0, m.f ? m.f[3 * i + 1] : 0.0, m.f ? m.f[3 * i + 2] : 0.0, m.drho ? m.drho[i] : 0.0);
  a.cgm[i] = make_double4(m.cg ? m.cg[3 * i] : 0.0, m.cg ? m.cg[3 * i + 1] : 0.0, m.cg ? m.cg[3 * i + 2] : 0.0, ms);
  a.e[i] = m.e ? m.e[i] : 0.0; a.de[i] = m.de ? m.de[i] : 0.0; a.cv[i] = m.cv ? m.cv[i] : 0.0;
  a.tag[i] = m.tag ? m.tag[i] : i + 1; a.mask[i] = m.mask ? m.mask[i] : 1; a.orig[i] = i;
}
// scatter back to LAMMPS local order (orig)
__global__ void k_unpack_atoms(int n, HostMirror m, PackArrays a, int multiphase, const int *outpos)
{
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = outpos ? outpos[s] : a.orig[s];
  double4 x = a.xt[s], vr = a.vr[s], v = a.vm[s], f = a.fd[s];
  if (m.x) { m.x[3 * i] = x.x; m.x[3 * i + 1] = x.y; m.x[3 * i + 2] = x.z; }
  if (m.v) { m.v[3 * i] = v.x; m.v[3 * i + 1] = v.y; m.v[3 * i + 2] = v.z; }
  if (m.vest) { m.vest[3 * i] = vr.x; m.vest[3 * i + 1] = vr.y; m.vest[3 * i + 2] = vr.z; }
  if (m.f) { m.f[3 * i] = f.x; m.f[3 * i + 1] = f.y; m.f[3 * i + 2] = f.z; }
  if (m.cg) { double4 c = multiphase ? a.cgm[s] : make_double4(0, 0, 0, 0); m.cg[3 * i] = c.x; m.cg[3 * i + 1] = c.y; m.cg[3 * i + 2] = c.z; }
  if (m.rho) m.rho[i] = vr.w;
  if (m.drho) m.drho[i] = f.w;
  if (m.e) m.e[i] = a.e[s];
  if (m.de) m.de[i] = a.de[s];
  if (m.cv) m.cv[i] = a.cv[s];
  if (m.rmass) m.rmass[i] = v.w;
  if (m.type) m.type[i] = tw_type(__double_as_longlong(x.w));
  if (m.mask) m.mask[i] = a.mask[s];
  if (m.tag) m.tag[i] = a.tag[s];
}

__global__ void k_fill_int(int n, int *p, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
