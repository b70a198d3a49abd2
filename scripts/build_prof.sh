#!/bin/bash
# Builds sgdnet_b200/libsgdnet_b200_prof.so: the library with the wavefront kernel's cycle counters compiled in
# (-DSGD_WAVE_PROF), read by scripts/wave_prof.py. Not part of the product build.
set -e
cd "$(dirname "$0")/.."
./build.sh > /dev/null
cd sgdnet_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
nvcc $FLAGS -DSGD_WAVE_PROF ${EXTRA:-} -c saga_sparse.cu -o ../../build/saga_sparse_prof.o 2>&1 | grep -i -A3 error || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsgdnet_b200_prof.so ../../build/saga_dense.o ../../build/saga_sparse_prof.o ../../build/passes.o ../../build/host_setup.o ../../build/engine.o -lcudart
echo built prof
