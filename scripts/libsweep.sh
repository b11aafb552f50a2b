#!/bin/bash
# development: compare library builds (mvdet_b200/libmvwarp_<name>.so, selected with MVWARP_LIB) on one workload
for lib in mvdet_b200/libmvwarp*.so; do
  for bv in ${BWD_VARIANTS:-0}; do
    echo -n "$(basename $lib) BWD_VARIANT=$bv "; MVWARP_LIB=$PWD/$lib MVWARP_BWD_VARIANT=$bv bash scripts/sweep.sh "$@" | sed 's/.*:: //'
  done
done
