#!/bin/bash
# Tuning aid: one short bench.py run, condensed to a line.
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | \
  python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g' % d['value'], 'step %.3f ms' % d['ms_per_step'], 'solver %.3f' % d['roofline']['ms_per_launch'], 'pyr %.3f' % d['roofline_pyramid']['ms_per_launch'], 'pyr_frac %.3f' % d['roofline_pyramid']['frac'], 'e2e %.4g' % d['e2e']['value'])"
