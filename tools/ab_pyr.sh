#!/bin/bash
# Tuning aid: pyramid time of every library in build_variants/ against the in-tree one (one batch in flight).
for i in 1 2; do
  for lib in lego_slam_b200/liblego_klt.so build_variants/*.so; do
    LEGO_KLT_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --streams 1 --no-side --no-sustained --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'pyr %.4f ms' % d['roofline_pyramid']['ms_per_launch'], 'frac %.3f' % d['roofline_pyramid']['frac'], 'step %.3f' % d['ms_per_step'])"
  done
done
