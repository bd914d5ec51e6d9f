#!/usr/bin/env bash
# Build libb200snark.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libb200snark.so
OBJ=../../build/obj
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
pids=()
for f in api group ntt dntt msm msm_acc_g1 msm_acc_g2 r1cs lcmap groth16 setup setup_groth16 serialize testops poly; do
  if [ ! -f "$OBJ/$f.o" ] || [ -n "$(find . ../../include -newer "$OBJ/$f.o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    ( s=$SECONDS; nvcc $FLAGS ${B2S_NVCC_EXTRA:-} -c -o "$OBJ/$f.o" "$f.cu"; echo "$f.cu: $((SECONDS-s))s" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/{api,group,ntt,dntt,msm,msm_acc_g1,msm_acc_g2,r1cs,lcmap,groth16,setup,setup_groth16,serialize,testops,poly}.o -lcudart_static -ldl -lrt -lpthread
echo "built $(readlink -f $OUT)"
