#!/bin/bash
# Last gpurun call of round 2 (~4 GPU-minutes left): the default bench line of the final tree (with the new `phases`
# object and the silent fraction), then three quick device-resident A/B lines at 1,024 streams:
# K0' alone on its SMs, the same with the tcgen05 recurrent core, and the tcgen05 core alone.
TAG=${1:-last}
mkdir -p gpurun_out/$TAG
timeout 170 python bench.py --steps 3 --warmup 3 > gpurun_out/$TAG/bench.json 2> gpurun_out/$TAG/bench.err; echo "bench rc=$?"
tail -2 gpurun_out/$TAG/bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$TAG/bench.json'))
    print('VALUE', round(d['value']), 'e2e', round(d['e2e']['value']), 'pcm16', round(d['e2e_pcm16']['value']), 'cpu', round(d['cpu_baseline']['value']), d['parity_vs_oracle'])
    print('PHASES', json.dumps(d['phases'])[:1500])
except Exception as e: print('bench parse failed', e)
PY
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-front-end --parity-streams 0"
for v in "CRISPY_NS_HP_EXCLUSIVE=1 CRISPY_NS_HP_PAR=1" "CRISPY_NS_HP_EXCLUSIVE=1 CRISPY_NS_HP_PAR=1 CRISPY_NS_RNN=tc5" "CRISPY_NS_RNN=tc5"; do
  n=$(echo "$v" | tr ' =' '__')
  env $v timeout 40 python bench.py $Q > gpurun_out/$TAG/ab_$n.json 2> gpurun_out/$TAG/ab_$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$TAG/ab_$n.json')); print('AB $v', round(d['value']), round(d['ms_per_step'],2))
except Exception as e: print('AB $v failed', e)
PY
done
