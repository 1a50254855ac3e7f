#!/bin/bash
# Build worm_algorithm_b200/libworm_b200_<name>.so with extra -D flags for ONE source file (A/B experiments;
# select at run time with WORM_B200_SO=<path>).
#   tools/build_variant.sh fold "-DWORM_PACK_FOLD=1"                 (worm_pack.cu, the default)
#   SRC=worm_lat tools/build_variant.sh prof "-DWORM_LAT_PROF=1"
set -e
name=$1; shift
src=${SRC:-worm_pack}
cd "$(dirname "$0")/../worm_algorithm_b200/csrc"
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr"
$NV $@ -c $src.cu -o /tmp/${src}_$name.o
objs=""
for o in c_api worm_mt worm_philox worm_lat worm_wide worm_pack postproc pca; do
  if [ $o = $src ]; then objs="$objs /tmp/${src}_$name.o"; else objs="$objs $o.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libworm_b200_$name.so $objs -cudart shared
echo built ../libworm_b200_$name.so
