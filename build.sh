#!/bin/bash
# Builds, in-tree and for sm_100a only:
#   qrochet.jl_b200/lib/libqrochet_b200.so       the product (include/qrochet_b200.h)
#   qrochet.jl_b200/lib/libqrochet_b200_diag.so  micro-benchmarks / hardware probes (include/qrochet_b200_diag.h)
#   tests/abi_smoke                              a plain-C program that drives the header through a C compiler
set -e
cd "$(dirname "$0")"
SRC=qrochet.jl_b200/csrc
OUT=qrochet.jl_b200/lib
mkdir -p $OUT build build/diag
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC"
stale() {  # $1 source, $2 object
  [ ! -f $2 ] || [ $1 -nt $2 ] || [ -n "$(find $SRC include \( -name '*.cuh' -o -name '*.h' \) -newer $2)" ]
}
objs=""; dobjs=""; pids=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if stale $f $o; then nvcc $FLAGS -c $f -o $o & pids="$pids $!"; fi
done
for f in $SRC/diag/*.cu; do
  o=build/diag/$(basename ${f%.cu}).o
  dobjs="$dobjs $o"
  if stale $f $o; then nvcc $FLAGS -c $f -o $o & pids="$pids $!"; fi
done
for p in $pids; do wait $p; done
# objects of sources that no longer exist must not be linked
for o in build/*.o; do
  [ -f $SRC/$(basename ${o%.o}).cu ] || rm -f $o
done
nvcc -shared $ARCH -o $OUT/libqrochet_b200.so $objs -lcudart -ldl
nvcc -shared $ARCH -o $OUT/libqrochet_b200_diag.so $dobjs -L$OUT -lqrochet_b200 -lcudart -Xlinker -rpath -Xlinker '$ORIGIN'
if [ -f tests/abi_smoke.c ]; then
  gcc -std=c99 -Wall -Werror -O1 -Iinclude tests/abi_smoke.c -o tests/abi_smoke -L$OUT -lqrochet_b200 \
      -Wl,-rpath,'$ORIGIN/../qrochet.jl_b200/lib' -lm
fi
echo "built $OUT/libqrochet_b200.so $OUT/libqrochet_b200_diag.so"
