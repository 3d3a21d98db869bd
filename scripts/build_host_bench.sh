#!/bin/bash
# builds scripts/host_solver_bench (CPU only) into gpurun_out/ or /tmp
set -e
cd "$(dirname "$0")/.."
O=${1:-/tmp/host_solver_bench}
D=$(mktemp -d)
for f in solver host_factors marginalize; do g++ -O3 -std=c++17 -fPIC -Iglio_b200/csrc -c glio_b200/csrc/$f.cpp -o $D/$f.o; done
g++ -O3 -mavx2 -mfma -std=c++17 -Iglio_b200/csrc -c glio_b200/csrc/band_chol_avx2.cpp -o $D/bc.o
g++ -O2 -std=c++17 -Iglio_b200/csrc scripts/host_solver_bench.cpp $D/*.o -o $O
echo $O
