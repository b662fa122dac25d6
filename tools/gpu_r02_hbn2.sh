# second (and last) short validation of the round: elapsed time under fix dt/reset on the device, the water_collapse thermo idiom and
# two more decks through lmp_b200 (shell changes: hybrid map, gravity / dt-reset scalars), tile path still the default
mkdir -p gpurun_out/hbn
(timeout 50 python -m pytest -m gpu -q --timeout 40 -p no:cacheprovider tests/test_gpu_parity.py::test_elapsed_time_under_fix_dt_reset_matches_oracle tests/test_gpu_lammps_shell.py::test_water_collapse_thermo_keywords "tests/test_gpu_lammps_shell.py::test_same_deck_reference_vs_b200[dam2d_dtreset-30-1e-09]" "tests/test_gpu_lammps_shell.py::test_same_deck_reference_vs_b200[droplet3d-10-1e-09]" tests/test_gpu_lammps_shell.py::test_phase_change_state_survives_a_second_run tests/test_gpu_tile.py::test_tile_path_is_the_one_that_runs > gpurun_out/hbn/shell.log 2>&1; echo "rc=$?" >> gpurun_out/hbn/shell.log)
tail -30 gpurun_out/hbn/shell.log | cut -c1-300
