# USER-B200: install (1), update (2) or remove (0) the /b200 shells in a LAMMPS source tree.
#
#   cp -r <repo>/lammps-sph-multiphase_b200/lammps/USER-B200 <lammps>/src/
#   cd <lammps>/src && make yes-user-sph yes-user-b200 && make B200_HOME=<repo> <machine>
#
# src/Makefile runs this script from inside src/USER-B200 (its yes-% / no-% rules, src/Makefile:208-236).
# B200_HOME is the checkout that holds include/b200_sph.h and lammps-sph-multiphase_b200/csrc/libb200sph.so
# (built by `python __graft_entry__.py build`); Makefile.b200sph turns it into include / library / rpath flags (it can also be given on the make command line).
# The shells derive from the USER-SPH classes, so nothing is copied unless USER-SPH (multiphase fork) is installed.

mode=$1
need=pair_sph_taitwater_multiphase.cpp     # a USER-SPH file only the multiphase fork has
tag=b200sph

if (test "$mode" != 0 && test ! -e ../$need) then
  echo "  USER-B200 needs USER-SPH first (make yes-user-sph): ../$need not found"
  exit 1
fi

for f in b200_shell.h pair_sph_b200.h pair_sph_b200.cpp fix_b200.h fix_b200.cpp verlet_b200.h verlet_b200.cpp; do
  if (test "$mode" = 0) then
    rm -f ../$f
  elif (! cmp -s $f ../$f) then
    cp $f ..
    if (test "$mode" = 2) then echo "  updating src/$f"; fi
  fi
done

# Makefile.package gets three make variables (every word mentions b200sph, so removing them is one sed); their values live in
# Makefile.b200sph next to this script, which Makefile.package.settings includes (paths there are relative to src/Obj_<machine>)
if (test -e ../Makefile.package) then
  sed -i -e "s|[^ \t]*$tag[^ \t]* ||g" ../Makefile.package
  if (test "$mode" != 0) then
    sed -i -e "s|^PKG_INC =[ \t]*|&\$(${tag}_INC) |" ../Makefile.package
    sed -i -e "s|^PKG_PATH =[ \t]*|&\$(${tag}_PATH) |" ../Makefile.package
    sed -i -e "s|^PKG_LIB =[ \t]*|&\$(${tag}_LIB) |" ../Makefile.package
  fi
fi
if (test -e ../Makefile.package.settings) then
  sed -i -e "/^include.*$tag.*\$/d" ../Makefile.package.settings
  if (test "$mode" != 0) then
    echo "include ../USER-B200/Makefile.$tag" >> ../Makefile.package.settings
  fi
fi
