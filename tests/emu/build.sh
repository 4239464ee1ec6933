#!/bin/sh
# TEST TOOLING: builds the CPU warp-emulator harness of the kernel bodies.
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -O2 -g -std=c++17 -fPIC -shared -Wall -Wno-unused-function -Wno-unknown-pragmas \
    -o _build/libemu_kernels.so emu_kernels.cpp simt_emu.cpp
