#!/bin/bash
# Build worm_algorithm_b200/libworm_b200_<name>.so with extra -D flags for worm_pack.cu (A/B experiments;
# select at run time with WORM_B200_SO=<path>).   tools/build_variant.sh fold "-DWORM_PACK_FOLD=1"
set -e
name=$1; shift
cd "$(dirname "$0")/../worm_algorithm_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr"
$NV $@ -c worm_pack.cu -o /tmp/worm_pack_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libworm_b200_$name.so c_api.o worm_mt.o worm_philox.o worm_lat.o worm_wide.o /tmp/worm_pack_$name.o postproc.o pca.o -cudart shared
echo built ../libworm_b200_$name.so
