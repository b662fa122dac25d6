# multiphase build after the pipelined exact tests: source-level capture; A/B of the build's CTA size on the C3 styles
mkdir -p gpurun_out/r02o
for nt in 128 256; do echo "== B200_BUILD_NT=$nt"; B200_BUILD_NT=$nt timeout 300 python tests/dev_bench.py c3 100 20 2>&1 | grep -E "ms/step|neigh_build|force " | cut -c1-160; done
S="python tests/dev_bench.py c3 100 3"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_build' -s 4 -c 1 -o gpurun_out/r02o/mpbuild -f $S > gpurun_out/r02o/ncu1.log 2>&1; echo "ncu mp build rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tile_force_mp' -s 2 -c 1 -o gpurun_out/r02o/mpforce -f $S > gpurun_out/r02o/ncu2.log 2>&1; echo "ncu mp force rc=$?"
