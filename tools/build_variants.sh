#!/bin/bash
# Builds tuning variants of the library side by side: fenics_b200/libfenics_b200_<tag>.so (select with FX_LIB=<path>).
# usage: tools/build_variants.sh tag1="-DFLAG=1 ..." tag2="..."    (the default build is restored at the end)
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag="${spec%%=*}"; flags="${spec#*=}"
  FX_NVCC_EXTRA="$flags" python -m fenics_b200.build --force > /dev/null
  cp fenics_b200/libfenics_b200.so "fenics_b200/libfenics_b200_${tag}.so"
  echo "built $tag ($flags)"
done
python -m fenics_b200.build --force > /dev/null
echo "default build restored"
