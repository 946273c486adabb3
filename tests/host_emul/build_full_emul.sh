#!/bin/sh
# Builds tests/host_emul/_build/libchannel_b200_emul.so: the library's own sources compiled with g++, kernels run by
# the CTA emulator (cta_emul.hpp), the CUDA runtime replaced by fake_cudart.cpp.  Test infrastructure only.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../channel_b200/csrc
OUT=$HERE/_build
mkdir -p $OUT/full
FLAGS="-O1 -std=c++17 -fPIC -w -pthread -ffp-contract=off -fvisibility=default -I/usr/local/cuda/include -I$HERE -include $HERE/emul_prelude.hpp"
for f in chb_api conv_kernels layout_kernels bodyforce_kernels rhs_kernel solve_kernels xpass3_kernels zpass3_kernels transpose restart_io convvel green_ctx; do
  g++ $FLAGS -x c++ -c $SRC/$f.cu -o $OUT/full/$f.o &
done
g++ -O2 -std=c++17 -fPIC -w -ffp-contract=off -c $SRC/host_tables.cpp -o $OUT/full/host_tables.o &
g++ -O1 -std=c++17 -fPIC -w -I/usr/local/cuda/include -c $HERE/fake_cudart.cpp -o $OUT/full/fake_cudart.o &
wait
g++ -shared -pthread -Wl,-Bsymbolic -o $OUT/libchannel_b200_emul.so $OUT/full/*.o -ldl
echo built $OUT/libchannel_b200_emul.so
