/* sph_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * The oracle exports the same entry points as include/b200_sph.h under the
 * prefix osph_ (so one Python driver exercises both through the same calls).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#define b200_sph               osph_sph
#define b200_create            osph_create
#define b200_destroy           osph_destroy
#define b200_last_error        osph_last_error
#define b200_version           osph_version
#define b200_comm_unique_id    osph_comm_unique_id
#define b200_comm_init         osph_comm_init
#define b200_domain            osph_domain
#define b200_atom_style        osph_atom_style
#define b200_boundary          osph_boundary
#define b200_get_box           osph_get_box
#define b200_neighbor          osph_neighbor
#define b200_timestep          osph_timestep
#define b200_comm_modify       osph_comm_modify
#define b200_atom_modify       osph_atom_modify
#define b200_pair_clear        osph_pair_clear
#define b200_pair_add          osph_pair_add
#define b200_fix_clear         osph_fix_clear
#define b200_fix_meso          osph_fix_meso
#define b200_fix_meso_stationary osph_fix_meso_stationary
#define b200_fix_gravity       osph_fix_gravity
#define b200_fix_phase_change  osph_fix_phase_change
#define b200_fix_setmeso       osph_fix_setmeso
#define b200_fix_enforce2d     osph_fix_enforce2d
#define b200_fix_setmeso_var   osph_fix_setmeso_var
#define b200_fix_addforce      osph_fix_addforce
#define b200_formula_check     osph_formula_check
#define b200_fix_setforce      osph_fix_setforce
#define b200_fix_setmesode     osph_fix_setmesode
#define b200_fix_dt_reset      osph_fix_dt_reset
#define b200_get_timestep      osph_get_timestep
#define b200_set_time          osph_set_time
#define b200_get_time          osph_get_time
#define b200_request_virial    osph_request_virial
#define b200_get_virial        osph_get_virial
#define b200_set_atoms         osph_set_atoms
#define b200_get_natoms        osph_get_natoms
#define b200_get_atoms         osph_get_atoms
#define b200_setup             osph_setup
#define b200_run               osph_run
#define b200_initial_integrate osph_initial_integrate
#define b200_final_integrate   osph_final_integrate
#define b200_neigh_decide      osph_neigh_decide
#define b200_forward_comm      osph_forward_comm
#define b200_reneighbor        osph_reneighbor
#define b200_force_clear       osph_force_clear
#define b200_pair_compute      osph_pair_compute
#define b200_pair_compute_all  osph_pair_compute_all
#define b200_reverse_comm      osph_reverse_comm
#define b200_post_force        osph_post_force
#define b200_get_neighbor_list osph_get_neighbor_list
#define b200_get_counters      osph_get_counters
#define b200_set_timing        osph_set_timing
#define b200_get_timers        osph_get_timers
#define b200_timer_name        osph_timer_name
#define b200_sync              osph_sync
#define b200_pair_desc         osph_pair_desc
#define b200_phase_change_desc osph_phase_change_desc
#define b200_atoms             osph_atoms

#include "../include/b200_sph.h"

/* oracle-only extras (ghost view, for debugging the parity traps) */
#ifdef __cplusplus
extern "C" {
#endif
/* copy owned+ghost arrays in the oracle's own (= reference) order */
int osph_get_all(osph_sph *h, int nmax, osph_atoms *a);
/* P ranks emulated in one process: ranks[r] was given osph_comm_init(n, r, procgrid, myloc, procneigh) before osph_domain.
 * Verlet::setup / Verlet::run in lock step, CommBrick's collectives between the ranks' arrays (sph_oracle.c "P ranks in one process") */
int osph_world_setup(osph_sph **ranks, int n);
int osph_world_run(osph_sph **ranks, int n, int nsteps);
#ifdef __cplusplus
}
#endif
#endif
