#!/bin/bash
# 8-GPU round, kept short (charged 8x): single-process e2e over every GPU + copy ceiling, torchrun bench.
TAG=${1:-multi8}; NG=${2:-8}
mkdir -p gpurun_out/$TAG; nvidia-smi topo -m > gpurun_out/$TAG/topo.txt 2>&1; nproc >> gpurun_out/$TAG/topo.txt; free -g >> gpurun_out/$TAG/topo.txt; numactl -H >> gpurun_out/$TAG/topo.txt 2>&1
timeout 400 python scripts/multi_e2e.py 1024 20 > gpurun_out/$TAG/multi_e2e.json 2> gpurun_out/$TAG/multi_e2e.err; tail -c 1200 gpurun_out/$TAG/multi_e2e.json; tail -3 gpurun_out/$TAG/multi_e2e.err
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/$TAG/bench.json 2> gpurun_out/$TAG/bench.err; tail -4 gpurun_out/$TAG/bench.err
TAG=$TAG python - <<'PY'
import json, os
for l in open('gpurun_out/%s/bench.json' % os.environ['TAG']):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('e2e_pcm16') or {}).get('value'))
PY
