#!/bin/bash
# Tuning aid: headline value (two batches in flight) and single-stream step of every library in build_variants/.
for i in 1 2; do
  for lib in lego_slam_b200/liblego_klt.so build_variants/*.so; do
    for s in 1 2; do
      LEGO_KLT_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --streams $s --no-side --no-sustained --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib streams $s', '%.3f ms' % d['ms_per_step'], '%.4g' % d['value'], 'solver %.3f' % d['roofline']['ms_per_launch'])"
    done
  done
done
