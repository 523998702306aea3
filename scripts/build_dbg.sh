#!/bin/bash
# Debug build of the library (clock64/printf instrumentation, -DW3_DEBUG) next to the
# release one; scripts/dbg.py loads it.  Never benchmark this build.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall \
  -DW3_DEBUG $W3_DBG_FLAGS -shared -o tools/libw3codec_dbg.so weath3rb0i_b200/csrc/w3_kernels.cu weath3rb0i_b200/csrc/w3_capi.cu
