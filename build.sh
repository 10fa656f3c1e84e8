#!/bin/bash
# Builds libqrochet_b200.so (sm_100a only) in-tree.
set -e
cd "$(dirname "$0")"
SRC=qrochet.jl_b200/csrc
OUT=qrochet.jl_b200/lib
mkdir -p $OUT build
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
objs=""
pids=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find $SRC include -name '*.cuh' -newer $o -o -name '*.h' -newer $o)" ]; then
    nvcc $FLAGS -c $f -o $o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libqrochet_b200.so $objs -lcudart -ldl
echo "built $OUT/libqrochet_b200.so"
