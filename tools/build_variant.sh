#!/bin/bash
# A second build of the library with extra nvcc flags (timing ablations, role timers), next to the shipped one:
#   tools/build_variant.sh ablate -DVIDX_TC_ABLATE   ->  vector-indexer_b200/lib_ablate/libvidx_b200.so  (VIDX_B200_LIB selects it)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../vector-indexer_b200"
NVCC=/usr/local/cuda/bin/nvcc
mkdir -p build_$NAME lib_$NAME
for f in search_kernels scan_tc kmeans_kernels kmeans_host index persist comm; do
  $NVCC -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -fmad=false --cudart static \
     -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wno-unknown-pragmas "$@" -c csrc/$f.cu -o build_$NAME/$f.o &
done
/usr/bin/g++ -std=c++17 -O3 -fPIC -ffp-contract=off -fno-fast-math -Wno-psabi -c csrc/chacha_blocks.cpp -o build_$NAME/chacha_blocks.o
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ --cudart static -shared -o lib_$NAME/libvidx_b200.so build_$NAME/*.o -lpthread -ldl -lrt
ls -la lib_$NAME/libvidx_b200.so
