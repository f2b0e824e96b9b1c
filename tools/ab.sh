#!/bin/bash
# Tuning aid: A/B of every library in build_variants/ against the in-tree one on ONE box: solver / pyramid time per step
# (one batch in flight, no side records), integer and sub-pixel keypoints, two rounds.   bash tools/ab.sh [PAIRS]
for i in 1 2; do
  for lib in lego_slam_b200/liblego_klt.so build_variants/*.so; do
    echo -n "$lib: "
    PAIRS=${1:-256} LEGO_KLT_LIB=$PWD/$lib python tools/subpixel_stats.py 2>&1 | tail -2 | awk '{printf "%s sol %s  ", $1, $5}'; echo
  done
done
