#!/bin/bash
# Ablation builds (tools/ablate.py times them; never shipped): libidgb200_ablateN.so = the library with
# the kernels named in KERNELS (default: the pipelined row-column degridder's file) compiled with -DIDGB200_ABLATE=N,
# every other object as the Makefile built it (untuned objects: N = 0 is the like-for-like baseline).
#   bit 0 (1)  no MMAs                                   (gridder_tc.cu, degridder_tc8.cu, gridder_sep.cu, degridder_sep.cu pipeline)
#   bit 1 (2)  operand stores predicated off at run time (gridder_tc.cu, degridder_tc8.cu)
#   bit 2 (4)  consumers skip the sum over the rows      (degridder_sep.cu pipeline); producers skip the A rows (gridder_sep.cu)
#   bit 3 (8)  producers skip the A rows                 (degridder_sep.cu pipeline); the B rows (gridder_sep.cu)
#   bit 4 (16) setup warps skip the B operand            (degridder_sep.cu pipeline)
set -e
cd "$(dirname "$0")/.."
C=ska_sdp_idg_bench_b200/csrc
make -C $C >/dev/null
mkdir -p tools/bin
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$C --expt-relaxed-constexpr"
ALL="gridder degridder gridder_tc gridder_sep degridder_sep scratch degridder_tc degridder_tc8 adder subgrid_fft"
KERNELS=${KERNELS:-degridder_sep}
OTHERS=""
for k in $ALL; do
  case " $KERNELS " in *" $k "*) ;; *) OTHERS="$OTHERS $C/$k.untuned.o";; esac
done
for n in ${ABLATIONS:-0 1 4 8 12 13 16}; do
  for k in $KERNELS; do
    $NVCC $FLAGS -DIDGB200_ABLATE=$n -c -o tools/bin/$k.ablate$n.o $C/$k.cu &
  done
done
wait
for n in ${ABLATIONS:-0 1 4 8 12 13 16}; do
  OBJ=""
  for k in $KERNELS; do OBJ="$OBJ tools/bin/$k.ablate$n.o"; done
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libidgb200_ablate$n.so $OBJ $OTHERS $C/capi.o -lcudart -ldl
done
ls -la tools/bin/libidgb200_ablate*.so
