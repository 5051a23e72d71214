#!/bin/bash
# Developer tool: builds variants of libb200ca with -D switches of lenia_march.cu (only that object is recompiled) into
# pyca_b200/variants/<name>.so.  Usage: tools/march_ab.sh name "-DMARCH_PACKED_MUL=0 ..." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../pyca_b200/csrc"
make -j8 > /dev/null
mkdir -p ../variants build_ab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $flags \
       -c lenia_march.cu -o build_ab/lenia_march_$name.o 2> build_ab/$name.log
  grep -E "registers|spill" build_ab/$name.log | sed -n 1,4p | tr '\n' ' '; echo " <- $name"
  objs=$(ls build/*.o | grep -v lenia_march.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$name.so $objs build_ab/lenia_march_$name.o -cudart static
done
