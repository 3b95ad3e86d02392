#!/bin/bash
# usage: tools/ab_step.sh <views> <iters> <workload>  -- stage times of the default library and of every tools/variant_*.so
cd "$(dirname "$0")/.."
for lib in default tools/variant_*.so; do
  [ -e "$lib" ] || [ "$lib" = default ] || continue
  if [ "$lib" = default ]; then unset B2R_LIB; else export B2R_LIB=$PWD/$lib; fi
  echo "== $lib"
  timeout 300 python tools/profile_step.py $1 $2 $3 2>&1 | tail -3
done
