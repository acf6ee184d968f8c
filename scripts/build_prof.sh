#!/bin/sh
# Builds the -DCIP_POTRF_PROF=${PROF_LEVEL:-1} variant of the library (clock64 phase marks in potrf_diag_kernel) and the
# FP64 latency micro-benchmark into conicip.jl_b200/csrc/build_prof/ (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/../conicip.jl_b200/csrc"
make -j8 >/dev/null
mkdir -p build_prof
A="-gencode arch=compute_100a,code=sm_100a"
nvcc $A -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DCIP_POTRF_PROF=${PROF_LEVEL:-1} -c chol.cu -o build_prof/chol.o
nvcc $A -shared -o build_prof/libprof.so build_prof/chol.o build/gemm_nt.o build/layout_matvec.o build/cones.o build/sdp.o \
  build/peaks.o build/engine.o build/ipm.o build/preprocess.o build/nccl_dl.o -ldl -cudart static
nvcc $A -O3 -o build_prof/fp64_latency ../../scripts/fp64_latency.cu
