#!/bin/bash
# Builds liblpf_b200.so (host mini-FEM + sm_100a kernels + C-ABI) in-tree.  The apply kernels are compiled once per order
# (csrc/apply_order.cu with -DLPF_ORDER=p) in parallel.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXX=/usr/bin/g++
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin $CXX --expt-relaxed-constexpr --extended-lambda -Wno-deprecated-gpu-targets ${LPF_DEBUG_ENV:+-DLPF_DEBUG_ENV}"
ORDERS="1 2 3 4 5 6 7 8 9 10"
mkdir -p build
newer() { [ ! -e "$1" ] || [ -n "$(find csrc host ../include build.sh -newer "$1" \( -name '*.cu' -o -name '*.cuh' -o -name '*.hpp' -o -name '*.h' -o -name '*.cpp' -o -name build.sh \) -print -quit)" ]; }
pids=()
for p in $ORDERS; do
  if newer build/apply_p$p.o; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -DLPF_ORDER=$p -c csrc/apply_order.cu -o build/apply_p$p.o & pids+=($!)
  fi
done
if newer build/lpf_device.o; then $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c csrc/lpf_device.cu -o build/lpf_device.o & pids+=($!); fi
$CXX -O2 -std=c++17 -fPIC -I/usr/local/cuda/include -c csrc/lpf_comm.cpp -o build/lpf_comm.o
$CXX -O2 -std=c++17 -fPIC -c host/lpf_host.cpp -o build/lpf_host.o
$CXX -O2 -std=c++17 -fPIC -c host/lpf_host_capi.cpp -o build/lpf_host_capi.o
for pid in "${pids[@]}"; do wait $pid; done
OBJS="build/lpf_device.o build/lpf_comm.o build/lpf_host.o build/lpf_host_capi.o"
for p in $ORDERS; do OBJS="$OBJS build/apply_p$p.o"; done
$NVCC -shared -ccbin $CXX -o liblpf_b200.so $OBJS -lcudart -ldl
echo "built $(pwd)/liblpf_b200.so"
