#!/bin/bash
# tools/build_variant.sh NAME FILE.cu [-DFLAG ...]: ab/libsegma_NAME.so = the product objects with segma_b200/csrc/FILE.cu
# recompiled with the given flags -- for alternately timed A/Bs of one kernel (tools/ab_logmel.py, tools/ab_l0.py take
# any number of library paths).  The product library itself is never touched.
set -e
NAME=$1; SRC=$2; shift 2
python -m segma_b200.build > /dev/null
mkdir -p ab
STEM=$(basename "$SRC" .cu)
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -I include -c segma_b200/csrc/$STEM.cu -o ab/${STEM}_$NAME.o
OBJS=$(ls segma_b200/build/*.o | grep -v "/$STEM.o")
nvcc -shared -o ab/libsegma_$NAME.so $OBJS ab/${STEM}_$NAME.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built ab/libsegma_$NAME.so
