#!/bin/bash
# usage: tools/build_variant.sh <name> <nvcc flags...>   -> tools/variant_<name>.so (tuning build, git-ignored, travels with gpurun)
cd "$(dirname "$0")/.."
name=$1; shift
B2R_LIB_OUT=$PWD/tools/variant_$name.so B2R_NVCC_FLAGS="$*" python -m py_numpy_renderer_b200.build --force
