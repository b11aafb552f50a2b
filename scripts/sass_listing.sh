#!/bin/bash
# SASS listings of the production kernels (fp32 instantiations the Wildtrack bench launches) -> profiles/
OUT=${1:-profiles}
LIB=mvdet_b200/libmvwarp.so
for k in "fwd_cl_kernelIfLb1ELi2ELi4ELb1ELb0" "bwd_cl_kernelIfLi3ELb1ELb1ELb1" "fwd_generic_kernelIfLb1" "bwd_generic_kernelIf"; do
  name=$(echo $k | sed 's/If.*//')
  cuobjdump -sass -fun "$(cuobjdump -elf $LIB | grep -o "_ZN3mvw[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)" $LIB | grep -v "^\s*/\* 0x[0-9a-f]* \*/\s*$" | grep -v "^Fatbin\|^=====\|^arch =\|^code version\|^host =\|^compile_size\|^identifier\|^$" > $OUT/r01_sass_${name}_f32.txt
  echo "$name: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $OUT/r01_sass_${name}_f32.txt) instructions, LDG.E.128: $(grep -c 'LDG.E.128' $OUT/r01_sass_${name}_f32.txt), STG.*128: $(grep -c 'STG.E.*128' $OUT/r01_sass_${name}_f32.txt), FFMA: $(grep -c FFMA $OUT/r01_sass_${name}_f32.txt), SHFL: $(grep -c SHFL $OUT/r01_sass_${name}_f32.txt)"
done
# tensor-core / TMA kernels
for k in warp_conv1_direct_kernel warp_conv1_pipe_kernelILi1ELb1 conv3x3_tc2_kernel conv3x3_tc_kernelILi1 conv3x3_head_kernelILi2; do
  name=$(echo $k | sed 's/IL.*//')
  cuobjdump -sass -fun "$(cuobjdump -elf $LIB | grep -o "_ZN3mvw[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)" $LIB | grep -v "^\s*/\* 0x[0-9a-f]* \*/\s*$" | grep -v "^Fatbin\|^=====\|^arch =\|^code version\|^host =\|^compile_size\|^identifier\|^$" > $OUT/r01_sass_${name}.txt
  echo "$name: UTCHMMA $(grep -c UTCHMMA $OUT/r01_sass_${name}.txt) (2CTA: $(grep -c "UTCHMMA.2CTA" $OUT/r01_sass_${name}.txt)), UTMALDG $(grep -c UTMALDG $OUT/r01_sass_${name}.txt), UBLKCP $(grep -c UBLKCP $OUT/r01_sass_${name}.txt), LDTM $(grep -c LDTM $OUT/r01_sass_${name}.txt), UTCBAR $(grep -c UTCBAR $OUT/r01_sass_${name}.txt)"
done
