for lib in lego_slam_b200/liblego_klt.so build_variants/*.so; do
  LEGO_KLT_LIB=$PWD/$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:template -c 3 --csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline 2>/dev/null | grep template | awk -F, -v l=$lib '{print l, $NF}'
done
