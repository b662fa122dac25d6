# launch list of the multiphase step (C3 styles, 1 M particles): which kernels make up neigh_bin_sort_ghost
mkdir -p gpurun_out/r02m
S="python tests/dev_bench.py c3 100 2"
timeout 300 $S > gpurun_out/r02m/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02m/launches_c3.csv $S > gpurun_out/r02m/ncu.log 2>&1; echo "rc=$?"
S="python tests/dev_bench.py c4 100 2"
timeout 300 $S > gpurun_out/r02m/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/r02m/launches_c4.csv $S > gpurun_out/r02m/ncu4.log 2>&1; echo "rc=$?"
