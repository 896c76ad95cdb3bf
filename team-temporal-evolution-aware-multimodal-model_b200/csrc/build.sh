#!/bin/bash
# Build libteam_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libteam_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p build
objs=""
pids=""
for f in *.cu; do
  o=build/${f%.cu}.o
  objs="$objs $o"
  stale=0
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ ../../include/team_b200.h -nt "$o" ]; then stale=1; fi
  for h in *.cuh; do if [ "$h" -nt "$o" ]; then stale=1; fi; done
  if [ $stale = 1 ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
# shared cudart: only the runtime symbols actually used are imported (the static runtime carries every entry point)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared --cudart=shared -Xlinker -rpath=/usr/local/cuda/lib64 -o $OUT $objs -lcuda
echo "built $(realpath $OUT)"
