# source-level captures: multiphase build (C3 styles, 1 M particles), multiphase force
mkdir -p gpurun_out/r02k
S="python tests/dev_bench.py c3 100 3"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_build' -s 4 -c 2 -o gpurun_out/r02k/mpbuild -f $S > gpurun_out/r02k/ncu1.log 2>&1; echo "ncu mp build rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_force_mp' -s 2 -c 1 -o gpurun_out/r02k/mpforce -f $S > gpurun_out/r02k/ncu2.log 2>&1; echo "ncu mp force rc=$?"
ls -la gpurun_out/r02k
