#!/bin/bash
# Build an experimental variant of libpcc_b200.so: tools/build_variant.sh NAME "-DFLAG ..."  ->  _lib/variants/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=pointcloudcounterfactual_b200/_lib/variants; mkdir -p $out/obj_$name
objs=""
for f in $(python -c "from pointcloudcounterfactual_b200.build import SOURCES; print(' '.join(s[:-3] for s in SOURCES))"); do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
       -c pointcloudcounterfactual_b200/csrc/$f.cu -o $out/obj_$name/$f.o &
  objs="$objs $out/obj_$name/$f.o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/$name.so $objs
rm -rf $out/obj_$name
echo $out/$name.so
