#!/bin/sh
# TEST INFRASTRUCTURE ONLY: single-thread host emulation of the kernel bodies (-DWFB_EMU, see csrc/wfb_rt.h), used to
# debug kernel logic in a container without a GPU. Never shipped, never loaded by the product path.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/emu/_build
g++ -O2 -g -std=c++17 -DWFB_EMU $EMU_DEFS -fPIC -Wall -Wno-unused-function -Wno-unused-variable -Wno-maybe-uninitialized \
    -Wno-unknown-pragmas -Wno-unused-but-set-variable -Iinclude \
    -x c++ wfmash_b200/csrc/wfa_host.cu -x c++ wfmash_b200/csrc/sketch.cu -x c++ wfmash_b200/csrc/minmer_host.cu \
    -x c++ wfmash_b200/csrc/epilogue.cu -x c++ wfmash_b200/csrc/index_host.cu -x c++ wfmash_b200/csrc/chain_host.cu -shared -o tests/emu/_build/libwfb_emu.so
echo tests/emu/_build/libwfb_emu.so
