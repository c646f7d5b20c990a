#!/bin/bash
# development build of the library with -DLPV_H8_PHASE_TIMING (see tools/h8_phase_timing.py); output tools/ab/liblpvmpc_timing.so
cd "$(dirname "$0")/.."; mkdir -p tools/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp -shared \
  -DLPV_H8_PHASE_TIMING -I include -o tools/ab/liblpvmpc_timing.so autonomous-racing-lpv-mpp-mpc_b200/csrc/lpvmpc.cu
