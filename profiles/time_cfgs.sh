#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in 512,32 512,16 256,32 256,16; do
  echo "== cfg $cfg"; RB200_FORCE_CFG=$cfg timeout 120 python profiles/time_kernels.py 2>&1 | grep -E "K2"
done
