#!/bin/sh
# Builds the parity oracle (test infrastructure) in two flavours:
#   liboracle_det.so — transcendentals via oracle/detmath.h (bit-matchable by the GPU path)
#   liboracle_sys.so — transcendentals via glibc (what the Rust reference would call)
#   liboracle_fast.so — the TIMED CPU arm of bench.py: glibc libm, -O3 -march=x86-64-v3 (AVX2-class code on any current
#                        server CPU; not -march=native: the library is built in one container and timed on another box)
# -ffp-contract=off: Rust/LLVM never fuses a*b+c; -fno-fast-math: IEEE evaluation order.
set -e
cd "$(dirname "$0")"
mkdir -p _build
FLAGS="-O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC -Wall -Wno-unused-function -Wno-comment"
g++ $FLAGS oracle.cpp -o _build/liboracle_det.so
g++ $FLAGS -DORC_SYSTEM_LIBM oracle.cpp -o _build/liboracle_sys.so
g++ $FLAGS -O3 -march=x86-64-v3 -DORC_SYSTEM_LIBM oracle.cpp -o _build/liboracle_fast.so
echo "oracle built"
