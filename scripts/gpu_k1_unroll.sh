#!/bin/bash
# K1 inner-product unroll variants (NS_DOT_UNROLL), serialised per-kernel times.
TAG=${1:-k1unroll}
mkdir -p gpurun_out/$TAG /tmp/variants
build() { nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 -o /tmp/variants/$1.so crispy_b200/csrc/crispy_ns.cu crispy_b200/csrc/ns_host.cpp > gpurun_out/$TAG/build_$1.log 2>&1 || echo "build $1 failed"; }
build u2 "-DNS_DOT_UNROLL=2" & build u4 "" & build u8 "-DNS_DOT_UNROLL=8" & wait
for v in u4 u2 u8 u4; do
  CRISPY_NS_SERIAL=1 CRISPY_NS_LIB=/tmp/variants/$v.so timeout 200 python scripts/prof_kernels.py 1024 256 2>&1 | grep -E "pitch_kernel" | sed "s/^/$v serial: /"
  CRISPY_NS_LIB=/tmp/variants/$v.so timeout 200 python scripts/prof_kernels.py 1024 512 2>&1 | grep -E "step" | sed "s/^/$v pipeline: /"
done | tee gpurun_out/$TAG/results.txt
