#!/bin/bash
# Kernel experiments: build libtfb_b200 with extra -D flags for tfb_channels_packed.cu only
#   scripts/build_variant.sh <name> [-DTFB_PK_...=v ...]  ->  topoflow36_b200/_variants/libtfb_<name>.so
# (select it with TFB_LIB_PATH=...; the other objects must exist: run csrc/build.sh first)
set -euo pipefail
root="$(cd "$(dirname "$0")/.." && pwd)"
src="$root/topoflow36_b200/csrc"
name="$1"; shift
mkdir -p "$root/topoflow36_b200/_variants"
F=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true
   -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xcompiler -O2)
/usr/local/cuda/bin/nvcc "${F[@]}" "$@" -c "$src/tfb_channels_packed.cu" -o "/tmp/pk_$name.o"
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a \
  -o "$root/topoflow36_b200/_variants/libtfb_$name.so" "$src/tfb_capi.o" "$src/tfb_channels.o" \
  "/tmp/pk_$name.o" "$src/tfb_channels_ext.o" "$src/tfb_d8.o" "$src/tfb_sources.o" "$src/tfb_prep.o" -lcudart
echo "built $name"
