#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): launch list of the bench command, one full capture of the three
# hot kernels of an LM iteration at the headline batch size.  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_r2_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'scan_match|factor_pair|window_kernel' -s 6 -c 3 \
    -o gpurun_out/prof_r2 -f python scripts/profile_step.py --windows 4736 > gpurun_out/prof_r2.log 2>&1
cp 2dliw-slam_b200/csrc/liblvio2d.so gpurun_out/prof_r2_liblvio2d.so
ls -la gpurun_out | tail -8
