#!/bin/bash
# usage: build_variant.sh <name> "<extra nvcc flags>"  -> scratch/libmdx_<name>.so (A/B builds, selected with MDX_LIB_PATH)
set -e
name=$1; flags=$2
src=/root/repo/marl-delivery_b200/csrc
top=/tmp/mdx_var_$name
tmp=$top/pkg/csrc
rm -rf $top; mkdir -p $tmp
cp $src/*.cu $src/*.cuh $src/*.h $src/*.inl $src/*.cpp $src/build.sh $tmp/
cp -r /root/repo/include $top/include
( cd $tmp && MDX_NVCC_EXTRA="$flags" bash build.sh > /dev/null 2>&1 ) || { grep -E "error" $tmp/build.log | head; exit 1; }
cp $tmp/libmdx.so /root/repo/scratch/libmdx_$name.so
grep -A2 "ILi3ELi32ELi1ELb1ELi1E" $tmp/build.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name"
