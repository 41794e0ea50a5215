#!/bin/bash
# Ablation builds of the two default tensor-core kernels (tools/ablate.py times them; never shipped):
# libidgb200_ablateN.so = the library with gridder_tc.cu / degridder_tc8.cu compiled with
# -DIDGB200_ABLATE=N (tc_common.cuh: bit 0 no MMAs, bit 1 operand stores predicated off), every
# other object as the Makefile built it (untuned objects: N = 0 is the like-for-like baseline).
set -e
cd "$(dirname "$0")/.."
C=ska_sdp_idg_bench_b200/csrc
make -C $C >/dev/null
mkdir -p tools/bin
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$C --expt-relaxed-constexpr"
OTHERS=""
for k in gridder degridder gridder_tc8 gridder_tc4 degridder_tc gridder_tc3 adder subgrid_fft; do OTHERS="$OTHERS $C/$k.untuned.o"; done
for n in ${ABLATIONS:-0 1 2 3}; do
  for k in gridder_tc degridder_tc8; do
    $NVCC $FLAGS -DIDGB200_ABLATE=$n -c -o tools/bin/$k.ablate$n.o $C/$k.cu &
  done
done
wait
for n in ${ABLATIONS:-0 1 2 3}; do
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libidgb200_ablate$n.so \
    tools/bin/gridder_tc.ablate$n.o tools/bin/degridder_tc8.ablate$n.o $OTHERS $C/capi.o -lcudart -ldl
done
ls -la tools/bin/libidgb200_ablate*.so
