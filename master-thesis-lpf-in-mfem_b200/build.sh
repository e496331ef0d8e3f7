#!/bin/bash
# Builds liblpf_b200.so (host mini-FEM + sm_100a kernels + C-ABI) in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXX=/usr/bin/g++
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin $CXX --expt-relaxed-constexpr --extended-lambda"
mkdir -p build
$NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c csrc/lpf_device.cu -o build/lpf_device.o
$CXX -O2 -std=c++17 -fPIC -I/usr/local/cuda/include -c csrc/lpf_comm.cpp -o build/lpf_comm.o
$CXX -O2 -std=c++17 -fPIC -c host/lpf_host.cpp -o build/lpf_host.o
$CXX -O2 -std=c++17 -fPIC -c host/lpf_host_capi.cpp -o build/lpf_host_capi.o
$NVCC -shared -ccbin $CXX -o liblpf_b200.so build/lpf_device.o build/lpf_comm.o build/lpf_host.o build/lpf_host_capi.o -lcudart -ldl
echo "built $(pwd)/liblpf_b200.so"
