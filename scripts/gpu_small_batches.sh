#!/bin/bash
# Small batches: the biquad kernel alone on its SMs (CRISPY_NS_HP_EXCLUSIVE=1) against sharing them (=0).
TAG=${1:-small}
mkdir -p gpurun_out/$TAG
for n in 128 256 512 768 1024; do
  for x in 0 1; do
    CRISPY_NS_HP_EXCLUSIVE=$x timeout 200 python scripts/prof_kernels.py $n 1920 2>&1 | grep -E "step|highpass" | tr '\n' ' ' | sed "s/^/streams=$n exclusive=$x: /"; echo
  done
done | tee gpurun_out/$TAG/results.txt
