#!/bin/bash
# Build a variant of libddmp_b200.so with extra nvcc flags into dual_dmp_b200/lib/variants/<name>.so (kernel A/B runs):
#   scripts/build_variant.sh seg512 -DDDMP_SEG_ROWS=512   then   DDMP_LIB_PATH=dual_dmp_b200/lib/variants/seg512.so python ...
set -e
name=$1; shift
cd "$(dirname "$0")/../dual_dmp_b200/csrc"
mkdir -p ../lib/variants ../../build/variants/$name
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
    --expt-relaxed-constexpr -DDDMP_WITH_TC "$@" -c $f -o ../../build/variants/$name/${f%.cu}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../lib/variants/$name.so ../../build/variants/$name/*.o -lcuda
echo built ../lib/variants/$name.so
