#!/bin/bash
# The round's last gpurun call: every GPU test, the bench line, configs[3] at one GPU's share of an 8-GPU job, K0's second
# form on small batches (+ its per-role cycle counts), the ncu launch list of the bench command, one full ncu capture of
# ns_highpass_par_kernel, the f64-side issue-rate micro-benchmark, the single-frame latency.
# usage: scripts/gpu_final.sh <tag>     (writes gpurun_out/<tag>_*)
TAG=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-900 gpurun_out/${TAG}_bench.json; echo
timeout 200 python bench.py --config c4 --total-streams 512 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_512.json 2> gpurun_out/${TAG}_bench_c4_512.err; echo "bench c4/512 rc=$?"
cut -c1-200 gpurun_out/${TAG}_bench_c4_512.json; echo
for n in 512 256; do timeout 120 python scripts/prof_kernels.py $n 1920 2>&1 | head -2; done | tee gpurun_out/${TAG}_small_batches.txt
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -DNS_HP_CLOCKS -o /tmp/libcrispy_ns_clk.so crispy_b200/csrc/crispy_ns.cu crispy_b200/csrc/ns_host.cpp 2>/dev/null
CRISPY_NS_HP_PAR=1 CRISPY_NS_LIB=/tmp/libcrispy_ns_clk.so CRISPY_NS_SERIAL=1 timeout 100 python scripts/prof_kernels.py 1024 32 2>&1 | grep "K0 warp" | sort | awk '!seen[$3]++' | tee gpurun_out/${TAG}_k0_clocks.txt
CRISPY_NS_HP_PAR=1 CRISPY_NS_SERIAL=1 timeout 100 python scripts/prof_kernels.py 1024 256 2>&1 | head -2 | tee -a gpurun_out/${TAG}_small_batches.txt
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_rate scripts/micro/fp64_rate.cu && /tmp/fp64_rate) > gpurun_out/${TAG}_fp64_rate.txt 2>&1; cat gpurun_out/${TAG}_fp64_rate.txt
timeout 120 python scripts/frame_latency.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_frame_latency.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ns_ -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --seconds 2.56 --no-e2e --no-cpu-baseline --parity-streams 0 > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ns_highpass_par -s 2 -c 1 -f -o gpurun_out/${TAG}_k0par \
  python scripts/prof_kernels.py 512 144 > gpurun_out/${TAG}_ncu_k0par.log 2>&1; echo "ncu k0par rc=$?"
ls -la gpurun_out | grep ${TAG}
