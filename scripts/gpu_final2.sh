#!/bin/bash
# After a change to K0's second form only: every GPU test, its small-batch numbers and per-role cycles, configs[3] at 512.
# (The 1,024-stream bench line does not run that kernel: scripts/gpu_final.sh covers it.)
TAG=${1:-final2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for n in 512 256; do timeout 120 python scripts/prof_kernels.py $n 1920 2>&1 | head -2; done | tee gpurun_out/${TAG}_small_batches.txt
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DNS_HP_CLOCKS -o /tmp/libcrispy_ns_clk.so crispy_b200/csrc/crispy_ns.cu crispy_b200/csrc/ns_host.cpp 2>/dev/null
CRISPY_NS_HP_PAR=1 CRISPY_NS_LIB=/tmp/libcrispy_ns_clk.so CRISPY_NS_SERIAL=1 timeout 100 python scripts/prof_kernels.py 1024 32 2>&1 | grep "K0 warp" | sort | awk '!seen[$3]++' | tee gpurun_out/${TAG}_k0_clocks.txt
CRISPY_NS_HP_PAR=1 CRISPY_NS_SERIAL=1 timeout 100 python scripts/prof_kernels.py 1024 256 2>&1 | head -2 | tee -a gpurun_out/${TAG}_small_batches.txt
timeout 200 python bench.py --config c4 --total-streams 512 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_512.json 2> gpurun_out/${TAG}_bench_c4_512.err; echo "bench c4/512 rc=$?"
cut -c1-200 gpurun_out/${TAG}_bench_c4_512.json; echo
