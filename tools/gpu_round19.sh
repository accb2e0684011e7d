#!/bin/bash
# C5 (F = 512, 1024 rays): ncu launch list of two training steps + one full-frame render chunk timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c5.csv python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 \
    --ncu-range 2 --pretrain 1500 --no-cpu-baseline > gpurun_out/launch_c5.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_c5.csv > gpurun_out/launches_c5.md 2>&1; head -40 gpurun_out/launches_c5.md
tail -2 gpurun_out/launch_c5.log
