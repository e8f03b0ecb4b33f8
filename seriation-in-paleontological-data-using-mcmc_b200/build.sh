#!/bin/sh
# Builds libseriation_b200.so (CUDA kernels + C ABI) and the `mcmc` CLI for sm_100a, in-tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
cd "$HERE/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CC=${CC:-gcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$CC -std=gnu99 -O2 -fPIC -Wall -ffp-contract=off -c ser_host.c -o ser_host.o
$NVCC $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off $NVCC_EXTRA -c ser_kernels.cu -o ser_kernels.o
$NVCC $ARCH -shared -o ../${SER_OUT:-libseriation_b200.so} ser_kernels.o ser_host.o -lm -ldl
if [ -f mcmc_main.c ]; then
  $CC -std=gnu99 -O2 -Wall -c mcmc_main.c -o mcmc_main.o
  $CC -o ../mcmc mcmc_main.o -L.. -lseriation_b200 -Wl,-rpath,'$ORIGIN' -lm
fi
echo "built $HERE/${SER_OUT:-libseriation_b200.so}"
