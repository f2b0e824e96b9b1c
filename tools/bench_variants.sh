#!/bin/bash
# Tuning aid: runs bench.py against every library in build_variants/ (built with LEGO_KLT_NVCC_DEFS).
for lib in lego_slam_b200/liblego_klt.so build_variants/*.so; do
  LEGO_KLT_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'value %.4g' % d['value'], 'step %.3f ms' % d['ms_per_step'], 'solver %.3f' % d['roofline']['ms_per_launch'], 'pyr %.3f' % d['roofline_pyramid']['ms_per_launch'], 'e2e %.4g' % d['e2e']['value'])"
done
