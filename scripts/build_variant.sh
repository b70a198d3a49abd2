#!/bin/bash
# Builds sgdnet_b200/libsgdnet_b200_<NAME>.so: the library with saga_sparse.cu recompiled with extra -D flags
# (measurement variants of the wavefront kernel; not part of the product build).
# Usage: scripts/build_variant.sh NAME "-DSGD_WAVE_TRACE"      (or -DSGD_NO_FAST_CONFLICT, ...)
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/.."
./build.sh > /dev/null
cd sgdnet_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
nvcc $FLAGS $EXTRA -c saga_sparse.cu -o ../../build/saga_sparse_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsgdnet_b200_$NAME.so ../../build/saga_dense.o ../../build/saga_dense_cluster.o ../../build/saga_dense_cluster_generic.o ../../build/saga_sparse_$NAME.o ../../build/saga_sparse_centred.o ../../build/passes.o ../../build/rng.o ../../build/setup.o ../../build/host_setup.o ../../build/engine.o -lcudart -ldl
echo built $NAME
