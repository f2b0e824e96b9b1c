#!/bin/bash
# One GPU-box visit: tests, default bench (own + reference arm), launch list and full ncu capture of one step.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
# profiling runs use ONE batch in flight, so that the launch order is ingest x2, then (l01, band, template, warp, lane<families>, lane) per step
timeout 900 python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline --no-sequence > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_v5.csv \
    python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline --no-sequence > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 20 -c 6 -o gpurun_out/prof_r01_v5 -f \
    python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline --no-sequence > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
