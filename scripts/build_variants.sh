#!/bin/bash
# Build A/B variants of the connector into ab/ (git-ignored; travels to the GPU box):
#   scripts/build_variants.sh name:"-DFLAG ..." [name:"flags" ...]
# Each variant is the whole library compiled with the extra flags; run them with QB_CUC_LIB=ab/liboctree_cuc_<name>.so
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/ab
pids=()
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  [ "$flags" = "$v" ] && flags=""
  ( make -C $ROOT/qubatron_b200/csrc -B OUT=$ROOT/ab/liboctree_cuc_$name.so EXTRA="$flags -Xptxas -v" > /tmp/build_$name.log 2>&1 \
      && grep -A2 "Compiling entry function '_ZN2qb18render_fast_kernelILi0ELb1ELb0ELb0E" /tmp/build_$name.log | tail -1 | sed "s/^/$name: /" \
      || { echo "$name: BUILD FAILED"; tail -5 /tmp/build_$name.log; } ) &
  pids+=($!)
done
wait
