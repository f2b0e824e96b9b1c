// triangulate_sm100.cu -- stereo / multi-view triangulation of tracked features, sm_100a.
//
// Replaces legoslam::triangulation (include/legoslam/algorithm.h:11-34), the step after
// Frontend::FindFeaturesInRight in Frontend::TriangulateNewPoints / BuildInitMap (src/frontend_g2o.cpp:111-155,
// :310-349): for every feature
//     A(2i,   :) = points[i][0] * m_i.row(2) - m_i.row(0)          m_i = poses[i].matrix3x4()
//     A(2i+1, :) = points[i][1] * m_i.row(2) - m_i.row(1)
//     pt_world   = (V.col(3) / V(3,3)).head<3>()   of the SVD A = U S V^T (singular values descending)
//     return finite(pt_world) && S[3] / S[2] < singRatioThr
// The reference computes the SVD with Eigen's bdcSvd (third party, not under /root/reference; for fewer than 16
// columns it is Eigen's two-sided JacobiSVD after a QR step).  Singular values and the null-space direction are
// unique, so any backward-stable SVD gives the same result to rounding; here: Givens QR of the 2n x 4 rows into a
// 4x4 R (same singular values, same V), then a one-sided (Hestenes) Jacobi SVD of R in fp64, which is accurate to
// high RELATIVE precision -- what the S[3]/S[2] test needs.  V.col(3)/V(3,3) is invariant to the sign of the column.
//
// One thread per feature; the camera poses are kernel arguments (shared by all features of a call, as in the
// callers: {camera_left->pose(), camera_right->pose()}).  This is fp64 arithmetic on ~40 bytes per feature: bound
// by the fp64 pipe, not by HBM; no tensor-core work.
#include "klt_kernels.h"

namespace legoklt {

namespace {

struct TriPoses {
    double m[kTriMaxViews][12];  // row-major 3x4 of every view
    int n_views;
};

struct TriCamera {
    double fx, fy, cx, cy;
};

// Rotates row `r` (4 entries, entries < k already zero) into the upper-triangular R.
__device__ __forceinline__ void givens_insert(double (&R)[4][4], double (&r)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double a = R[k][k], b = r[k];
        if (b != 0.0) {
            const double h = hypot(a, b);
            const double c = a / h, s = b / h;
#pragma unroll
            for (int j = k; j < 4; ++j) {
                const double t0 = R[k][j], t1 = r[j];
                R[k][j] = c * t0 + s * t1;
                r[j] = c * t1 - s * t0;
            }
        }
    }
}

// One-sided Jacobi SVD of the 4x4 G (columns are rotated until mutually orthogonal): on return the singular values
// are the column norms of G and V holds the right singular vectors (unsorted).
__device__ __forceinline__ void jacobi_svd4(double (&G)[4][4], double (&V)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    alpha += G[i][p] * G[i][p];
                    beta += G[i][q] * G[i][q];
                    gamma += G[i][p] * G[i][q];
                }
                if (gamma != 0.0 && fabs(gamma) > 1.0e-16 * sqrt(alpha * beta)) {
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double gp = G[i][p], gq = G[i][q];
                        G[i][p] = c * gp - s * gq;
                        G[i][q] = s * gp + c * gq;
                        const double vp = V[i][p], vq = V[i][q];
                        V[i][p] = c * vp - s * vq;
                        V[i][q] = s * vp + c * vq;
                    }
                }
            }
        if (!rotated) break;
    }
}

// algorithm.h:24-33 from R: world point and the reference's return value.
__device__ __forceinline__ bool finish(double (&R)[4][4], double thr, double (&pt)[3]) {
    double G[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[i][j] = (j >= i) ? R[i][j] : 0.0;
    jacobi_svd4(G, V);
    double sv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sv[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j] + G[3][j] * G[3][j]);
    // smallest and second smallest singular value (S[3], S[2] of the descending order)
    int i3 = 0;
    double smin = sv[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (sv[j] < smin) {
            smin = sv[j];
            i3 = j;
        }
    double s3 = 0, s2 = 1.7976931348623157e308, v0 = 0, v1 = 0, v2 = 0, v3 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j == i3) {
            s3 = sv[j];
            v0 = V[0][j], v1 = V[1][j], v2 = V[2][j], v3 = V[3][j];
        } else if (sv[j] < s2) {
            s2 = sv[j];
        }
    }
    pt[0] = v0 / v3;
    pt[1] = v1 / v3;
    pt[2] = v2 / v3;
    if (not_finite(pt[0]) || not_finite(pt[1]) || not_finite(pt[2])) return false;  // :27-29
    return s3 / s2 < thr;                                                            // :31-34
}

__device__ __forceinline__ void add_view(double (&R)[4][4], const double *m, double x, double y) {
    double r0[4], r1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        r0[j] = x * m[8 + j] - m[j];      // :20
        r1[j] = y * m[8 + j] - m[4 + j];  // :21
    }
    givens_insert(R, r0);
    givens_insert(R, r1);
}

// Generic: normalised image points (points[i][0], points[i][1] of the reference's VecVec3), n_views per feature.
__global__ void __launch_bounds__(128)
triangulate_kernel(const __grid_constant__ TriPoses poses, const double *__restrict__ points, int n, double thr,
                   double *__restrict__ pt_world, uint8_t *__restrict__ ok) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double R[4][4] = {};
    for (int v = 0; v < poses.n_views; ++v) {
        const double2 p = *reinterpret_cast<const double2 *>(points + ((size_t)f * poses.n_views + v) * 2);
        add_view(R, poses.m[v], p.x, p.y);
    }
    double pt[3];
    const bool good = finish(R, thr, pt);
    pt_world[3 * (size_t)f] = pt[0];
    pt_world[3 * (size_t)f + 1] = pt[1];
    pt_world[3 * (size_t)f + 2] = pt[2];
    ok[f] = good ? 1 : 0;
}

// Stereo, fused with Camera::pixel2camera (src/camera.cpp:21-25, depth 1): pixel keypoints as the tracker leaves
// them (packed float2) -> world points.  Features with valid[f] == 0 (tracking failed: the reference has no right
// feature for them, src/frontend_g2o.cpp:114-115) are skipped: ok = 0, point = 0.
__global__ void __launch_bounds__(128)
triangulate_stereo_kernel(const __grid_constant__ TriPoses poses, TriCamera cl, TriCamera cr,
                          const float2 *__restrict__ kp_left, const float2 *__restrict__ kp_right,
                          const uint8_t *__restrict__ valid, int n, double thr, double *__restrict__ pt_world,
                          uint8_t *__restrict__ ok) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double pt[3] = {0, 0, 0};
    bool good = false;
    if (!valid || valid[f]) {
        const float2 a = kp_left[f], b = kp_right[f];
        double R[4][4] = {};
        add_view(R, poses.m[0], ((double)a.x - cl.cx) * 1.0 / cl.fx, ((double)a.y - cl.cy) * 1.0 / cl.fy);
        add_view(R, poses.m[1], ((double)b.x - cr.cx) * 1.0 / cr.fx, ((double)b.y - cr.cy) * 1.0 / cr.fy);
        good = finish(R, thr, pt);
    }
    pt_world[3 * (size_t)f] = pt[0];
    pt_world[3 * (size_t)f + 1] = pt[1];
    pt_world[3 * (size_t)f + 2] = pt[2];
    ok[f] = good ? 1 : 0;
}

TriPoses make_poses(const double *poses34, int n_views) {
    TriPoses p;
    p.n_views = n_views;
    for (int v = 0; v < kTriMaxViews; ++v)
        for (int j = 0; j < 12; ++j) p.m[v][j] = (v < n_views) ? poses34[12 * v + j] : 0.0;
    return p;
}

}  // namespace

cudaError_t launch_triangulate(const double *poses34, int n_views, const double *d_points, int n, double thr,
                               double *d_pt_world, uint8_t *d_ok, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    triangulate_kernel<<<(n + 127) / 128, 128, 0, stream>>>(make_poses(poses34, n_views), d_points, n, thr, d_pt_world, d_ok);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_triangulate_stereo(const double *poses34, const double *cam_left, const double *cam_right,
                                      const float2 *d_kp_left, const float2 *d_kp_right, const uint8_t *d_valid, int n,
                                      double thr, double *d_pt_world, uint8_t *d_ok, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const TriCamera cl = {cam_left[0], cam_left[1], cam_left[2], cam_left[3]};
    const TriCamera cr = {cam_right[0], cam_right[1], cam_right[2], cam_right[3]};
    triangulate_stereo_kernel<<<(n + 127) / 128, 128, 0, stream>>>(make_poses(poses34, 2), cl, cr, d_kp_left, d_kp_right,
                                                                  d_valid, n, thr, d_pt_world, d_ok);
    note_launch();
    return cudaGetLastError();
}

}  // namespace legoklt
