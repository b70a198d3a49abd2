#!/bin/bash
# Builds libsgdnet_b200.so (sm_100a) in-tree. Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")/sgdnet_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
mkdir -p ../../build
pids=""
for f in saga_dense saga_dense_cluster saga_dense_cluster_generic saga_sparse saga_sparse_centred passes rng setup host_setup engine; do
  if [ ! -f ../../build/$f.o ] || [ $f.cu -nt ../../build/$f.o ] || [ common.cuh -nt ../../build/$f.o ] || [ kernels.h -nt ../../build/$f.o ] || [ host_setup.h -nt ../../build/$f.o ] || [ setup.h -nt ../../build/$f.o ] || [ ../../include/sgdnet_b200.h -nt ../../build/$f.o ]; then
    rm -f ../../build/$f.o     # a failed compile must not leave a stale object for the link below
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o ../../build/$f.o &
    pids="$pids $!"
  fi
done
for pid in $pids; do wait $pid || { echo "build.sh: a compile failed" >&2; exit 1; }; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libsgdnet_b200.so ../../build/saga_dense.o ../../build/saga_dense_cluster.o ../../build/saga_dense_cluster_generic.o ../../build/saga_sparse.o ../../build/saga_sparse_centred.o ../../build/passes.o ../../build/rng.o ../../build/setup.o ../../build/host_setup.o ../../build/engine.o -lcudart -ldl
echo built sgdnet_b200/libsgdnet_b200.so
