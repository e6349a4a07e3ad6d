#!/bin/bash
# e2e (host buffers, f32 and PCM16) as a function of the host path's copy granularity, on one box.
mkdir -p gpurun_out/$1
for ch in 32 64 128 256 512; do
  CRISPY_NS_HOST_CHUNK_FRAMES=$ch timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --parity-streams 0 > gpurun_out/$1/e2e_$ch.json 2> gpurun_out/$1/e2e_$ch.err
  python - <<PY
import json
d=json.load(open('gpurun_out/$1/e2e_$ch.json'))
print('host chunk $ch frames: value', round(d['value']), 'e2e f32', round(d['e2e']['value']), 'e2e pcm16', round(d['e2e_pcm16']['value']))
PY
done
