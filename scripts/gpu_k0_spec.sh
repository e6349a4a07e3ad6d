#!/bin/bash
# One gpurun call for the time-parallel biquad (K0, NS_HP_PAR): the GPU parity tests that finish in a minute (the long-run
# configs run in the round's final call), then the product library (NS_HP_PAR=1) against the single-recursion-warp build
# (build_variants/libcrispy_ns_par0.so, -DNS_HP_PAR=0), pipelined, serialised (isolated kernel times) and on small batches.
# usage: scripts/gpu_k0_spec.sh <tag>
TAG=${1:-k0par}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for v in par0 par1; do
  lib=$PWD/build_variants/libcrispy_ns_${v}.so; [ $v = par1 ] && lib=$PWD/crispy_b200/libcrispy_ns.so
  for n in 1024 512 256; do
    echo "=== $v, $n streams, pipelined"
    CRISPY_NS_LIB=$lib timeout 120 python scripts/prof_kernels.py $n 512 2>&1 | head -2 | tee -a gpurun_out/${TAG}_${v}_kernels.txt
  done
  echo "=== $v, 1024 streams, serialised (isolated kernel times)"
  CRISPY_NS_LIB=$lib CRISPY_NS_SERIAL=1 timeout 120 python scripts/prof_kernels.py 1024 256 2>&1 | head -3 | tee -a gpurun_out/${TAG}_${v}_kernels.txt
done
echo "=== par1, 1024 streams, K0 alone on its SMs"
CRISPY_NS_HP_EXCLUSIVE=1 timeout 120 python scripts/prof_kernels.py 1024 512 2>&1 | head -2 | tee -a gpurun_out/${TAG}_par1_kernels.txt
timeout 120 python scripts/frame_latency.py 2>&1 | tail -3 | tee gpurun_out/${TAG}_frame_latency.txt
