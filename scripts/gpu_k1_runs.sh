#!/bin/bash
# K1 frames-per-CTA variants (more, smaller CTAs per SM): per-kernel times in the pipeline and serialised.
TAG=${1:-k1runs}
mkdir -p gpurun_out/$TAG /tmp/variants
build() { nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 -o /tmp/variants/$1.so crispy_b200/csrc/crispy_ns.cu crispy_b200/csrc/ns_host.cpp > gpurun_out/$TAG/build_$1.log 2>&1 || echo "build $1 failed"; }
build base "" & build r6 "-DNS_PITCH_RUN=6" & build r5 "-DNS_PITCH_RUN=5" & build r6b "-DNS_PITCH_RUN=6 -DNS_PITCH_THREADS=288" & wait
run() { # name lib chunk
  for serial in 0 1; do
    CRISPY_NS_SERIAL=$serial CRISPY_NS_LIB=/tmp/variants/$2.so CRISPY_NS_CHUNK_FRAMES=$3 timeout 200 python scripts/prof_kernels.py 1024 480 2>&1 | grep -E "step|pitch_kernel" | sed "s/^/$1 chunk=$3 serial=$serial: /"
  done
}
{ run base base 32; run base base 30; run r6 r6 30; run r6 r6 36; run r5 r5 30; run r6b r6b 30; } | tee gpurun_out/$TAG/results.txt
CRISPY_NS_LIB=/tmp/variants/r6.so CRISPY_NS_CHUNK_FRAMES=30 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "batch_matches or chunk_size or golden" 2>&1 | tail -3 | tee gpurun_out/$TAG/pytest_r6.txt
