#!/bin/bash
# multi-GPU round: multi-device C-ABI test, single-process e2e over every GPU + copy ceiling, torchrun bench (both arms).
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-multi}; NG=${2:-2}
mkdir -p gpurun_out/$TAG; nvidia-smi topo -m > gpurun_out/$TAG/topo.txt 2>&1; nproc >> gpurun_out/$TAG/topo.txt; free -g >> gpurun_out/$TAG/topo.txt
timeout 300 python -m pytest tests/test_gpu_configs.py -q -k "multi_device" 2>&1 | tail -3 | tee gpurun_out/$TAG/pytest_multi.txt
timeout 600 python scripts/multi_e2e.py 1024 20 > gpurun_out/$TAG/multi_e2e.json 2> gpurun_out/$TAG/multi_e2e.err; tail -c 1500 gpurun_out/$TAG/multi_e2e.json; tail -3 gpurun_out/$TAG/multi_e2e.err
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $NG --steps 3 --warmup 3) > gpurun_out/$TAG/ref2.json 2> gpurun_out/$TAG/ref2.err; tail -4 gpurun_out/$TAG/ref2.err
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 3 --warmup 3) > gpurun_out/$TAG/bench2.json 2> gpurun_out/$TAG/bench2.err; tail -4 gpurun_out/$TAG/bench2.err
TAG=$TAG python - <<'PY'
import json
import os
for f in ('gpurun_out/%s/ref2.json' % os.environ['TAG'], 'gpurun_out/%s/bench2.json' % os.environ['TAG']):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('e2e_pcm16') or {}).get('value'), d.get('roofline',{}) and d['roofline'].get('kernel'))
PY
