#!/bin/bash
# Build tuning variants of libcrispy_ns.so (extra -D flags) next to the product library and time each on the GPU.
# usage (under gpurun): scripts/variants.sh <tag> "<name>:<nvcc flags>" ...
TAG=$1; shift
mkdir -p gpurun_out /tmp/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  so=/tmp/variants/libcrispy_ns_${name}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $flags \
     -o $so crispy_b200/csrc/crispy_ns.cu crispy_b200/csrc/ns_host.cpp > gpurun_out/${TAG}_${name}_build.log 2>&1 || { echo "build $name failed"; tail -5 gpurun_out/${TAG}_${name}_build.log; continue; }
  echo "=== variant $name ($flags)"
  CRISPY_NS_LIB=$so timeout 300 python scripts/prof_kernels.py 1024 512 2>&1 | tee gpurun_out/${TAG}_${name}_kernels.txt
  # isolated (serialised, cold) duration of each kernel of one chunk
  CRISPY_NS_LIB=$so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ns_ -s 14 -c 7 --csv \
     --log-file gpurun_out/${TAG}_${name}_iso.csv python scripts/prof_kernels.py 1024 64 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_${name}_iso.csv')) if len(r)>5 and r[0].isdigit()]
print('  isolated us:', ', '.join(f"{r[4].split('(')[0].replace('ns_','').replace('_kernel','')}={float(r[-1].replace(',',''))/1000:.0f}" for r in rows))
PY
done
if [ -n "$CHUNKS" ]; then
  for ch in $CHUNKS; do
    echo "=== product library, CRISPY_NS_CHUNK_FRAMES=$ch"
    CRISPY_NS_CHUNK_FRAMES=$ch timeout 300 python scripts/prof_kernels.py 1024 512 2>&1 | tee gpurun_out/${TAG}_chunk${ch}_kernels.txt
  done
fi
