#!/bin/bash
# Development tool: build mgf_b200/lib/libmgfb<NAME>.so from a git ref (or the working tree with "WT") plus extra nvcc flags,
# for tools/solver_ab.py A/B runs.   usage: tools/build_variant.sh NAME REF|WT [-DFLAG ...]
set -e
NAME=$1; REF=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT
if [ "$REF" != "WT" ]; then
  SRC=$(mktemp -d); git -C "$ROOT" archive "$REF" mgf_b200/csrc include | tar -x -C "$SRC"
fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o "$ROOT/mgf_b200/lib/libmgfb$NAME.so" "$SRC/mgf_b200/csrc/capi.cu"
echo "built libmgfb$NAME.so from $REF $*"
