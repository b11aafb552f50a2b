#!/bin/bash
# SASS listings of the production kernels (fp32 instantiations) -> profiles/
OUT=${1:-profiles}
LIB=mvdet_b200/libmvwarp.so
for k in "fwd_cl_kernelIfLb1ELi2ELi4ELb0" "bwd_cl_kernelIfLi3ELb1" "fwd_generic_kernelIfLb1" "bwd_generic_kernelIf"; do
  name=$(echo $k | sed 's/If.*//')
  cuobjdump -sass -fun "$(cuobjdump -elf $LIB | grep -o "_ZN3mvw[0-9]*${k}[A-Za-z0-9_]*" | sort -u | head -1)" $LIB | grep -v "^\s*/\* 0x[0-9a-f]* \*/\s*$" | grep -v "^Fatbin\|^=====\|^arch =\|^code version\|^host =\|^compile_size\|^identifier\|^$" > $OUT/r01_sass_${name}_f32.txt
  echo "$name: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $OUT/r01_sass_${name}_f32.txt) instructions, LDG.E.128: $(grep -c 'LDG.E.128' $OUT/r01_sass_${name}_f32.txt), STG.*128: $(grep -c 'STG.E.*128' $OUT/r01_sass_${name}_f32.txt), FFMA: $(grep -c FFMA $OUT/r01_sass_${name}_f32.txt), SHFL: $(grep -c SHFL $OUT/r01_sass_${name}_f32.txt)"
done
