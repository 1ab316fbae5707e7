#!/bin/bash
# Profiling variant of the library: the TD kernel built with its clock64 timeline hooks
# (RB200_TC_TIMELINE=1).  Use with RB200_LIB=reagent_b200/libreagent_b200_timeline.so.
set -euo pipefail
cd "$(dirname "$0")/../reagent_b200/csrc"
./build.sh >/dev/null
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets \
  -DRB200_TC_TIMELINE=1 "$@" -c rb200_dqn_tc.cu -o build/rb200_dqn_tc_timeline.o
objs=$(ls build/rb200_*.o | grep -v -e rb200_dqn_tc.o -e _timeline.o)
$NVCC -shared -o ../libreagent_b200_timeline.so $objs build/rb200_dqn_tc_timeline.o -lcudart
echo "built ../libreagent_b200_timeline.so"
