#!/bin/bash
# tools/build_logmel_variant.sh N: ab/libsegma_v$N.so = the product objects with logmel.cu compiled as -DSEGMA_LOGMEL_VARIANT=N
set -e
N=$1
python -m segma_b200.build > /dev/null
mkdir -p ab
nvcc -DSEGMA_LOGMEL_VARIANT=$N -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I include -c segma_b200/csrc/logmel.cu -o ab/logmel_v$N.o
OBJS=$(ls segma_b200/build/*.o | grep -v '/logmel.o')
nvcc -shared -o ab/libsegma_v$N.so $OBJS ab/logmel_v$N.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built ab/libsegma_v$N.so
