#!/bin/sh
# TEST INFRASTRUCTURE ONLY: single-thread host emulation of the kernel bodies (-DWFB_EMU, see csrc/wfb_rt.h), used to
# debug kernel logic in a container without a GPU. Never shipped, never loaded by the product path.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/emu/_build
OUT=tests/emu/_build/libwfb_emu.so
# up to date (and built with the same EMU_DEFS)? then do not pay the half minute again
if [ -f "$OUT" ] && [ "$(cat tests/emu/_build/defs 2>/dev/null)" = "$EMU_DEFS" ] && \
   [ -z "$(find wfmash_b200/csrc include tests/emu/build_emu.sh -newer "$OUT" -type f | head -1)" ]; then
  echo "$OUT"; exit 0
fi
g++ -O2 -g -std=c++17 -DWFB_EMU -DWFB_HOST_PAR_MIN_BYTES=0 $EMU_DEFS -fPIC -Wall -Wno-unused-function -Wno-unused-variable -Wno-maybe-uninitialized \
    -Wno-unknown-pragmas -Wno-unused-but-set-variable -Iinclude \
    -x c++ wfmash_b200/csrc/wfa_host.cu -x c++ wfmash_b200/csrc/sketch.cu -x c++ wfmash_b200/csrc/minmer_host.cu \
    -x c++ wfmash_b200/csrc/epilogue.cu -x c++ wfmash_b200/csrc/index_host.cu -x c++ wfmash_b200/csrc/chain_host.cu \
    -x c++ wfmash_b200/csrc/filter_host.cu -x c++ wfmash_b200/csrc/stats_host.cu -x c++ wfmash_b200/csrc/ani_host.cu -x c++ wfmash_b200/csrc/phases_host.cu -x c++ wfmash_b200/csrc/index_file_host.cu -shared -o tests/emu/_build/libwfb_emu.so
printf '%s' "$EMU_DEFS" > tests/emu/_build/defs
echo tests/emu/_build/libwfb_emu.so
