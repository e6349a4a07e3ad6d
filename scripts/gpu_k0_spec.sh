#!/bin/bash
# One gpurun call for the time-parallel biquad (K0): chain micro-benchmark, the GPU parity tests that finish in a minute
# (the long-run configs run in the round's final call), per-role cycle counts of the instrumented build (-DNS_HP_CLOCKS),
# and both forms of the kernel ($CRISPY_NS_HP_PAR = 0 / 1) pipelined at three batch sizes and serialised.
# usage: scripts/gpu_k0_spec.sh <tag>
TAG=${1:-k0par}
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hp_latency scripts/micro/hp_latency.cu 2>/dev/null && timeout 60 /tmp/hp_latency) 2>&1 | grep variant | tee gpurun_out/${TAG}_hp_latency.txt
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
CRISPY_NS_HP_PAR=1 CRISPY_NS_LIB=$PWD/build_variants/libcrispy_ns_clk.so CRISPY_NS_SERIAL=1 timeout 100 python scripts/prof_kernels.py 1024 32 2>&1 | grep "K0 warp" | sort | uniq -c | sort -k4n | awk '!seen[$4]++' | tee gpurun_out/${TAG}_k0_clocks.txt
for par in 0 1; do
  for n in 1024 768 512 256; do
    echo "=== CRISPY_NS_HP_PAR=$par, $n streams, pipelined"
    CRISPY_NS_HP_PAR=$par timeout 120 python scripts/prof_kernels.py $n 512 2>&1 | head -2 | tee -a gpurun_out/${TAG}_par${par}_kernels.txt
  done
  echo "=== CRISPY_NS_HP_PAR=$par, 1024 streams, serialised (isolated kernel times)"
  CRISPY_NS_HP_PAR=$par CRISPY_NS_SERIAL=1 timeout 120 python scripts/prof_kernels.py 1024 256 2>&1 | head -2 | tee -a gpurun_out/${TAG}_par${par}_kernels.txt
done
timeout 120 python scripts/frame_latency.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_frame_latency.txt
