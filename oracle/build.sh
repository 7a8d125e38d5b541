#!/bin/sh
# Builds the parity oracle (test infrastructure) in two flavours:
#   liboracle_det.so — transcendentals via oracle/detmath.h (bit-matchable by the GPU path)
#   liboracle_sys.so — transcendentals via glibc (what the Rust reference would call)
# -ffp-contract=off: Rust/LLVM never fuses a*b+c; -fno-fast-math: IEEE evaluation order.
set -e
cd "$(dirname "$0")"
mkdir -p _build
FLAGS="-O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC -Wall -Wno-unused-function -Wno-comment"
g++ $FLAGS oracle.cpp -o _build/liboracle_det.so
g++ $FLAGS -DORC_SYSTEM_LIBM oracle.cpp -o _build/liboracle_sys.so
echo "oracle built"
