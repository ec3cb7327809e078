#!/bin/bash
# Build libtfb_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libtfb_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
       -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
       -Xcompiler -fPIC -Xcompiler -O2)
objs=()
pids=()
for f in tfb_capi tfb_channels tfb_channels_packed tfb_channels_ext tfb_d8 tfb_sources tfb_prep; do
  if [ ! -f "$here/$f.o" ] || [ "$here/$f.cu" -nt "$here/$f.o" ] || \
     [ "$here/tfb_common.cuh" -nt "$here/$f.o" ] || [ "$here/tfb_channels_common.cuh" -nt "$here/$f.o" ] || [ "$here/tfb_pk.cuh" -nt "$here/$f.o" ] || [ "$here/../../include/tfb.h" -nt "$here/$f.o" ]; then
    rm -f "$here/$f.o"       # a failed compile must not leave a stale object to link
    "$NVCC" "${FLAGS[@]}" ${TFB_EXTRA_NVCC_FLAGS:-} -c "$here/$f.cu" -o "$here/$f.o" &
    pids+=($!)
  fi
  objs+=("$here/$f.o")
done
for pid in "${pids[@]}"; do
  wait "$pid" || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$out" "${objs[@]}" -lcudart
echo "built $out"
