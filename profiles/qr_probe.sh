#!/bin/bash
# QR-DQN (config 3) update: default vs tcgen05 weight gradients, then the launch list of one update
val() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'])"; }
python bench.py --config 3 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | val default
RB200_WGRAD_TC=1 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | val wgrad_tc
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_qr.csv \
  python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python profiles/launch_list.py gpurun_out/launches_qr.csv 45
