#!/bin/bash
# A/B builds of one kernel file with different -D knobs (tools/ablate.py --ab times them; never shipped):
#   KERNEL=degridder_sep tools/build_ab.sh name1:"-DX=1 -DY=2" name2:"-DX=0" ...
# -> tools/bin/libidgb200_ab_<name>.so, every other object as the Makefile built it (untuned objects).
set -e
cd "$(dirname "$0")/.."
C=ska_sdp_idg_bench_b200/csrc
make -C $C >/dev/null
mkdir -p tools/bin
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -I$C --expt-relaxed-constexpr"
ALL="gridder degridder gridder_tc gridder_sep degridder_sep scratch degridder_tc degridder_tc8 adder subgrid_fft"
KERNEL=${KERNEL:-degridder_sep}
OTHERS=""
for k in $ALL; do [ "$k" = "$KERNEL" ] || OTHERS="$OTHERS $C/$k.untuned.o"; done
rm -f tools/bin/libidgb200_ab_*.so
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  ( $NVCC $FLAGS $defs -c -o tools/bin/$KERNEL.ab_$name.o $C/$KERNEL.cu &&
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/libidgb200_ab_$name.so tools/bin/$KERNEL.ab_$name.o $OTHERS $C/capi.o -lcudart -ldl ) &
done
wait
ls tools/bin/libidgb200_ab_*.so
