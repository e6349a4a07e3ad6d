#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, per-kernel device times, ncu launch list + full capture.
# usage: scripts/gpu_round.sh <tag>      (writes gpurun_out/<tag>_*)
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 300 python scripts/prof_kernels.py 1024 512 > gpurun_out/${TAG}_kernels.txt 2>&1; cat gpurun_out/${TAG}_kernels.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ns_ -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --seconds 2.56 --no-e2e --no-cpu-baseline --parity-streams 0 > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns_ -s 14 -c 7 -f -o gpurun_out/${TAG}_full \
  python scripts/prof_kernels.py 1024 64 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hp_latency scripts/micro/hp_latency.cu && timeout 60 /tmp/hp_latency) > gpurun_out/${TAG}_hp_latency.txt 2>&1; cat gpurun_out/${TAG}_hp_latency.txt
ls -la gpurun_out | head -30
