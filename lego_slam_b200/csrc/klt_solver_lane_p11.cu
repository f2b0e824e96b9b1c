// klt_solver_lane_p11.cu -- the LANE solver (klt_solver_lane.cu) compiled for the 11x11 patch, offsets -5..5
// (BASELINE.json's stress configuration).  Same code, other compile-time shapes: seven sample pairs per grid row,
// 16 window rows, 11 template rows of 12 floats -- 1152 bytes of shared memory per thread, hence 64-thread CTAs
// (two to three per SM).
#define LANE_PATCH_LO (-5)
#define LANE_PATCH_HI 5
#define LANE_SUFFIX _p11
#define LANE_T 64
#define LANE_CTAS 3
#define LANE_TPL_T 64
#include "klt_solver_lane.cu"
