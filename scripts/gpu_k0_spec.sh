#!/bin/bash
# One gpurun call for the speculative biquad (K0, NS_HP_SPEC): chain latency micro-benchmark, the GPU parity tests that
# finish in a minute or two (the long-run configs run in the round's final call), and the three builds (prebuilt under
# build_variants/: -DNS_HP_SPEC=0 / 2; the product library is 1) timed in the pipeline, serialised, and on small batches.
# usage: scripts/gpu_k0_spec.sh <tag>
TAG=${1:-k0spec}
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hp_latency scripts/micro/hp_latency.cu 2>/dev/null && timeout 60 /tmp/hp_latency) > gpurun_out/${TAG}_hp_latency.txt 2>&1; grep variant gpurun_out/${TAG}_hp_latency.txt
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for v in spec0 spec1 spec2; do
  lib=$PWD/build_variants/libcrispy_ns_${v}.so; [ $v = spec1 ] && lib=$PWD/crispy_b200/libcrispy_ns.so
  for n in 1024 512 256; do
    echo "=== $v, $n streams, pipelined"
    CRISPY_NS_LIB=$lib timeout 120 python scripts/prof_kernels.py $n 512 2>&1 | head -2 | tee -a gpurun_out/${TAG}_${v}_kernels.txt
  done
  echo "=== $v, 1024 streams, serialised (isolated kernel times)"
  CRISPY_NS_LIB=$lib CRISPY_NS_SERIAL=1 timeout 120 python scripts/prof_kernels.py 1024 256 2>&1 | head -3 | tee -a gpurun_out/${TAG}_${v}_kernels.txt
done
