#!/bin/bash
# tools/make_cubin.sh <out.cubin> [-D...]   - the product kernels (csrc/pp_tuned.cu) as a post-processed cubin with extra
# macro definitions, e.g. -DO3D_PP_BODY_FILE='"/tmp/kv/anneal_101.inc"' for a candidate statement order. Development tool:
# microbench/kbench times such files against each other (KBENCH_CUBIN=a.cubin:b.cubin) before an order is adopted.
set -e
out=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I"$here/omega3d_b200/csrc" "$@" -cubin "$here/omega3d_b200/csrc/pp_tuned.cu" -o "$tmp/base.cubin"
python3 "$here/tools/sass_patch.py" "$tmp/base.cubin" "$tmp/stage.cubin" pp2_kernel > /dev/null
python3 "$here/tools/sass_patch.py" "$tmp/stage.cubin" "$out" ppc_kernel > /dev/null
rm -rf "$tmp"
