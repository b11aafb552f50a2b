#!/bin/bash
# SASS listings of the production kernels (cuobjdump works without a GPU) -> profiles/${TAG}_sass_*.txt
#   bash scripts/sass_listing.sh profiles r02
OUT=${1:-profiles}; TAG=${2:-r02}
LIB=mvdet_b200/libmvwarp.so
ELF=$(cuobjdump -elf $LIB)
dump() {  # mangled-name fragment, output name
  local sym=$(echo "$ELF" | grep -o "_ZN3mvw[0-9]*$1[A-Za-z0-9_]*" | sort -u | head -1)
  [ -z "$sym" ] && { echo "$2: kernel not found ($1)"; return; }
  cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -v "^\s*/\* 0x[0-9a-f]* \*/\s*$" | grep -v "^Fatbin\|^=====\|^arch =\|^code version\|^host =\|^compile_size\|^identifier\|^$" > $OUT/${TAG}_sass_$2.txt
  local f=$OUT/${TAG}_sass_$2.txt
  echo "$2: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $f) instr | LDG.E.128 $(grep -c 'LDG.E.128' $f) STG.*128 $(grep -c 'STG.E.*128' $f) FFMA $(grep -c FFMA $f) SHFL $(grep -c SHFL $f) ATOM/RED $(grep -c 'ATOM\|RED\.' $f) | UTCHMMA $(grep -c UTCHMMA $f) (2CTA $(grep -c 'UTCHMMA.2CTA' $f)) UTMALDG $(grep -c UTMALDG $f) UBLKCP $(grep -c UBLKCP $f) LDTM $(grep -c LDTM $f) UTCBAR $(grep -c UTCBAR $f) HMMA $(grep -c '[^C]HMMA' $f)"
}
dump "fwd_cl_kernelIfLb1ELi2ELi4ELb1ELb0" fwd_cl_kernel_f32
dump "bwd_cl_kernelIfLi3ELb1ELb1ELb1" bwd_cl_kernel_f32
dump "fwd_cl_kernelI13__nv_bfloat16Lb1ELi2ELi4ELb1ELb0" fwd_cl_kernel_bf16
dump "bwd_cl_kernelI13__nv_bfloat16Li4ELb1ELb1ELb1" bwd_cl_kernel_bf16
dump "fwd_nchw_kernelIfLb1" fwd_nchw_kernel_f32
dump "bwd_generic_kernelIf" bwd_generic_kernel_f32
dump "fwd_table_tile_kernelIfLi2ELi4ELb1" fwd_table_tile_kernel_f32
dump "transpose_tiles_kernelIf" transpose_tiles_kernel_f32
dump "copy_touched_kernelILi16" copy_touched_kernel
dump "relu_bwd_cast_kernel" relu_bwd_cast_kernel
dump "warp_conv1_direct_kernelILb0" warp_conv1_direct_kernel
dump "warp_conv1_direct_kernelILb1" warp_conv1_direct_kernel_ups
dump "conv3x3_tc2_kernel" conv3x3_tc2_kernel
dump "conv3x3_head_kernelILi2" conv3x3_head_kernel
dump "head_partial_kernel" head_partial_kernel
dump "head_stencil_kernel" head_stencil_kernel
