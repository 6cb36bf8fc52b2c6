#!/bin/bash
# Build tuning variants of the apply kernel for ONE order (only that order's object differs) and
# print ptxas register counts.  usage: tools/tune.sh D Q "tag -DB200PA_TUNE_NEB=.. -DB200PA_TUNE_MINB=.. ..." [...]
set -e
cd "$(dirname "$0")/../cardiac-ablation-ecm2_b200"
D=$1; Q=$2; shift 2
make -j8 >/dev/null
for cfg in "$@"; do
  set -- $cfg
  tag="d${D}_$1"; shift
  mkdir -p build_tune
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v \
     "$@" -DB200PA_D=$D -DB200PA_Q=$Q \
     -c csrc/elem_inst.cu -o build_tune/elem_$tag.o 2> build_tune/$tag.log
  echo "$tag: $(grep -A2 "pa_apply_kernelILi${D}ELi${Q}ELb1ELb1" build_tune/$tag.log | grep -E 'Used|spill' | tr '\n' ' ' | sed 's/ptxas info    ://; s/bytes//g')"
  objs=""
  for o in build/*.o; do
    if [ "$(basename $o)" == "elem_${D}_${Q}.o" ]; then objs="$objs build_tune/elem_$tag.o"; else objs="$objs $o"; fi
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libb200pa_$tag.so $objs -lcudart -ldl
done
