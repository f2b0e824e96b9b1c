// klt_solver_lane_inv.cu -- the LANE solver (klt_solver_lane.cu) compiled for the reference's INVERSE mode, 7x7 patch
// (BASELINE.json config C4: "inverse-compositional mode"; src/algorithm.cpp:57,59,74-87 -- including the stale
// Jacobian).  Three record planes per (feature, level) -- template, img1 x / y gradient differences --
// 1056 bytes of shared memory per thread, 64-thread CTAs, three per SM; the template kernel stages 44 KB per CTA.
#define LANE_PATCH_LO (-3)
#define LANE_PATCH_HI 3
#define LANE_SUFFIX _inv
#define LANE_INVERSE 1
#define LANE_T 64
#define LANE_CTAS 3
#define LANE_TPL_T 64
#include "klt_solver_lane.cu"
