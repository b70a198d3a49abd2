#!/bin/bash
# Builds sgdnet_b200/libsgdnet_b200_<NAME>.so with ONE source recompiled with extra -D flags (measurement variants).
# Usage: scripts/build_variant_file.sh NAME FILE(.cu without extension) "-DFLAG ..."
set -e
NAME=$1; FILE=$2; EXTRA=$3
cd "$(dirname "$0")/.."
./build.sh > /dev/null
cd sgdnet_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
nvcc $FLAGS $EXTRA -c $FILE.cu -o ../../build/${FILE}_$NAME.o
OBJS=""
for f in saga_dense saga_dense_cluster saga_dense_cluster_generic saga_sparse saga_sparse_centred passes rng setup host_setup engine; do
  if [ $f = $FILE ]; then OBJS="$OBJS ../../build/${FILE}_$NAME.o"; else OBJS="$OBJS ../../build/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsgdnet_b200_$NAME.so $OBJS -lcudart -ldl
echo built $NAME
