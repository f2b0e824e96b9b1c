// klt_solver_lane_p8.cu -- the LANE solver (klt_solver_lane.cu) compiled for the 8x8 patch, offsets -4..3:
// the patch BASELINE.json's metric text names ("4-level, 8x8 patch"; the reference itself computes 7x7,
// src/algorithm.cpp:40,63-64).  Same code, different compile-time patch bounds; the per-thread shared-memory
// footprint grows from 688 to 752 bytes (8 x 8-float template rows, 13 window rows); 96-thread CTAs, three per SM.
#define LANE_PATCH_LO (-4)
#define LANE_PATCH_HI 3
#define LANE_SUFFIX _p8
#define LANE_T 96
#define LANE_CTAS 3
#include "klt_solver_lane.cu"
