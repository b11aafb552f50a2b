#!/bin/bash
# Tuning knobs (MVWARP_*) exist in the development build only: exports MVWARP_LIB=mvdet_b200/libmvwarp_dev.so (python -m mvdet_b200.build --dev)
export MVWARP_LIB=${MVWARP_LIB:-$PWD/mvdet_b200/libmvwarp_dev.so}
# development: fused up-sample forward -- quad kernel (0, default) vs task kernel (1) vs tile kernel (2), fp32 / bf16
for dt in fp32 bf16; do
  for v in ${TBL_VARIANTS:-0 1 2}; do
    echo -n "dtype=$dt TBL_VARIANT=$v "
    MVWARP_TBL_VARIANT=$v bash scripts/sweep.sh "--fuse-upsample --dtype $dt $*" | sed 's/.*:: //'
  done
done
