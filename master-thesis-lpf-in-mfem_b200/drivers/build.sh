#!/bin/bash
# Builds the C++ drivers (twins of the reference's Solvers/ and Convergence_and_Scaling/ drivers) against
# liblpf_b200.so, plus the compile check of include/lpf_mfem_adapter.hpp against the stub mfem.hpp.
set -e
cd "$(dirname "$0")"
CXX=/usr/bin/g++
FLAGS="-O2 -std=c++17 -Wall -I../../include -pthread"
LINK="-L.. -llpf_b200 -Wl,-rpath,\$ORIGIN/../.. -pthread"
mkdir -p bin
for d in PF_linear_par_partial ss laplace_solver cylinder-diffraction convergence-parallel-partial; do
  $CXX $FLAGS $d.cpp $LINK -o bin/$d
done
$CXX $FLAGS -Istub adapter_check.cpp $LINK -o bin/adapter_check
echo "built drivers in $(pwd)/bin"
