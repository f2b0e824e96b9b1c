#!/bin/bash
# One GPU-box visit: tests, default bench (own + reference arm), launch list and full ncu capture of one step.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1
# profiling runs use ONE batch in flight and no side records, so that the launch order is ingest x2, then (l01, band, template,
# warp (deferred), lane<family>, lane) per device-resident step
PROF="python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline --no-side --no-sustained"
timeout 900 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv $PROF > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 20 -c 6 -o gpurun_out/prof_r02 -f $PROF > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
