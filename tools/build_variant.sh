#!/bin/bash
# build a tuning variant of libpnae.so into tools/wip/variants/libpnae_$1.so with extra nvcc defines $2
set -e
name=$1; defs=$2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/tools/wip/variants; tmp=$(mktemp -d)
mkdir -p $out
for f in $root/pointnet_autoencoder_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=true -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I $root/include $defs -c $f -o $tmp/$b.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libpnae_$name.so $tmp/*.o
rm -rf $tmp
echo built $out/libpnae_$name.so
