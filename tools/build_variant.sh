#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>  ->  mesoengine_b200/_variants/libmeso_<name>.so (git-ignored A/B builds)
set -e
cd "$(dirname "$0")/../mesoengine_b200/csrc"
name=$1; shift
mkdir -p ../_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  --shared -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 -cudart static "$@" \
  -o ../_variants/libmeso_$name.so meso_capi.cu k_voxelize.cu k_occupancy.cu k_raymarch.cu k_mesh.cu k_carve.cu k_resident.cu k_cubes.cu meso_group.cu
echo ../_variants/libmeso_$name.so
