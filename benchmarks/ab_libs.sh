#!/bin/bash
# A/B of several builds of libbsplat.so on one box: per-stage times of a frame for every library given
#   benchmarks/ab_libs.sh [config] lib1.so lib2.so ...      (paths relative to the repo root)
cfg=$1; shift
for l in "$@"; do
  echo "== $l"
  BSPLAT_LIB=$PWD/$l python benchmarks/stage_probe.py $cfg
done
