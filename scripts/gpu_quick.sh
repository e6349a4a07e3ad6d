#!/bin/bash
# Quick gpurun call: GPU parity tests, a bench line (no CPU baseline), per-kernel device times.
# usage: scripts/gpu_quick.sh <tag> [extra bench args]
TAG=${1:-quick}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('VALUE', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, 'parity', d['parity_vs_oracle'])
    for k in d['kernels']: print('  ', k['kernel'], round(k['avg_launch_us']), 'us', round(k['share_of_kernel_time'],3))
except Exception as e: print('bench parse failed', e)
PY
timeout 300 python scripts/prof_kernels.py 1024 512 > gpurun_out/${TAG}_kernels.txt 2>&1; cat gpurun_out/${TAG}_kernels.txt
