#!/bin/bash
# Build libvqw.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")/chainer_vq_vae_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
OBJS=""
for f in *.cu; do
  o="${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ mol.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/vqw.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS ${EXTRA_NVCC_FLAGS} -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -shared -o libvqw.so $OBJS -cudart static -lpthread -ldl
echo "built $(pwd)/libvqw.so"
