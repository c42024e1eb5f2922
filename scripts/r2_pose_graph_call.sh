#!/bin/bash
# First GPU call of round 2 for the pose-graph row: everything round 1 could not run.
#   gpurun --timeout 900 -- 'bash scripts/r2_pose_graph_call.sh'
# Writes under gpurun_out/: pg_check.txt (parity + wall clock, plain and partitioned paths), pg_tests.log (pytest, xfails
# run for real), pg_launches_plain.csv / pg_launches_segments.csv (ncu launch lists), pg_sanitizer.txt, pg_bench.json.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scripts/pg_gpu_check.py > gpurun_out/pg_check.log 2>&1
timeout 300 python -m pytest tests/test_zz_gpu_pose_graph.py tests/test_golden.py -m gpu -q --runxfail > gpurun_out/pg_tests.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pg_launches_plain.csv \
    python scripts/pg_gpu_profile.py > gpurun_out/pg_profile_plain.log 2>&1
LVIO2D_PG_SEGMENTS=64 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pg_launches_segments.csv \
    python scripts/pg_gpu_profile.py > gpurun_out/pg_profile_segments.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py > gpurun_out/pg_sanitizer.txt 2>&1
timeout 200 python scripts/pg_bench.py > gpurun_out/pg_bench.json 2> gpurun_out/pg_bench.err
tail -25 gpurun_out/pg_check.log; tail -5 gpurun_out/pg_tests.log; tail -3 gpurun_out/pg_sanitizer.txt
