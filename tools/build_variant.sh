#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=1 ..."  ->  fcl_b200/lib/variants/libfclgpu_NAME.so (A/B builds; not shipped)
set -e
cd "$(dirname "$0")/../fcl_b200/csrc"
mkdir -p ../lib/variants
nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas \
  -Wno-deprecated-gpu-targets $2 -shared -o ../lib/variants/libfclgpu_$1.so fclgpu_api.cu bvh_build.cpp comm.cpp mesh_io.cpp -ldl
