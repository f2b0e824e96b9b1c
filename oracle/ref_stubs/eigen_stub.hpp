// Stand-in for the part of Eigen the reference's hot path uses (TEST INFRASTRUCTURE ONLY -- oracle/_ref build).
//
// src/algorithm.cpp uses (line numbers of the reference file):
//   :55-57  Matrix2d::Zero(), Vector2d::Zero()            :68-80  -1.0 * Vector2d(a, b)
//   :83     b += -error * J                                :86     H += J * J.transpose()
//   :92-93  update = H.ldlt().solve(b)                     :94-95,107-108  update[0], update[1]
//   :113    update.norm()
// All of these are coefficient-wise double arithmetic in Eigen (no FMA without -mfma: the reference builds
// -std=c++11 -O3 only, CMakeLists.txt:6-7), written here coefficient by coefficient.
//
// THIRD PARTY, RESTATED: Eigen 3.3.x (un-pinned by the reference, CMakeLists.txt:15) Eigen/src/Cholesky/LDLT.h --
// ldlt_inplace<Lower>::unblocked() and LDLT::_solve_impl(), written for general n the way Eigen structures them
// (diagonal pivoting by largest |entry|, first on ties; "entire diagonal is zero" early exit; no scaling by an invalid
// pivot; solve = P^T L^-T D^+ L^-1 P with D^+ the pseudo-inverse at tolerance 1/highest()).  This is the one piece of
// the hot loop that is NOT the reference's own text in oracle/_ref.
#ifndef LEGO_REF_STUB_EIGEN_HPP
#define LEGO_REF_STUB_EIGEN_HPP

#include <cmath>
#include <limits>

namespace Eigen {

class RowVector2d;
class Matrix2d;

class Vector2d {
public:
    Vector2d() { v[0] = v[1] = 0; }
    Vector2d(double a, double b) { v[0] = a; v[1] = b; }
    static Vector2d Zero() { return Vector2d(0.0, 0.0); }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    Vector2d &operator+=(const Vector2d &o) { v[0] = v[0] + o.v[0]; v[1] = v[1] + o.v[1]; return *this; }
    inline RowVector2d transpose() const;
    double squaredNorm() const { return v[0] * v[0] + v[1] * v[1]; }
    double norm() const { return std::sqrt(squaredNorm()); }
    double v[2];
};

class RowVector2d {
public:
    double v[2];
};
inline RowVector2d Vector2d::transpose() const { RowVector2d r; r.v[0] = v[0]; r.v[1] = v[1]; return r; }

static inline Vector2d operator*(double s, const Vector2d &a) { return Vector2d(s * a[0], s * a[1]); }

namespace ref_stub {

// LDLT<Matrix<double,N,N>, Lower>: compute() then solve(), after Eigen 3.3 LDLT.h.
template <int N>
class LDLT {
public:
    explicit LDLT(const double (&a)[N][N]) {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) m[i][j] = a[i][j];
        compute();
    }
    template <typename Vec>
    Vec solve(const Vec &rhs) const {
        double dst[N];
        for (int i = 0; i < N; ++i) dst[i] = rhs[i];
        // dst = P b : transpositions applied in order
        for (int k = 0; k < N; ++k)
            if (tr[k] != k) { double t = dst[k]; dst[k] = dst[tr[k]]; dst[tr[k]] = t; }
        // dst = L^-1 (P b), L unit lower
        for (int i = 1; i < N; ++i) {
            double s = 0;
            for (int j = 0; j < i; ++j) s = (j == 0) ? m[i][j] * dst[j] : s + m[i][j] * dst[j];
            dst[i] = dst[i] - s;
        }
        // dst = D^+ (L^-1 P b): pseudo-inverse of D
        const double tolerance = 1.0 / std::numeric_limits<double>::max();
        for (int i = 0; i < N; ++i) {
            if (std::fabs(m[i][i]) > tolerance) dst[i] = dst[i] / m[i][i];
            else dst[i] = 0;
        }
        // dst = L^-T (...), unit upper = adjoint of L
        for (int i = N - 2; i >= 0; --i) {
            double s = 0;
            for (int j = i + 1; j < N; ++j) s = (j == i + 1) ? m[j][i] * dst[j] : s + m[j][i] * dst[j];
            dst[i] = dst[i] - s;
        }
        // dst = P^T (...): transpositions in reverse order
        for (int k = N - 1; k >= 0; --k)
            if (tr[k] != k) { double t = dst[k]; dst[k] = dst[tr[k]]; dst[tr[k]] = t; }
        Vec out;
        for (int i = 0; i < N; ++i) out[i] = dst[i];
        return out;
    }

private:
    void compute() {
        bool found_zero_pivot = false;
        (void)found_zero_pivot;
        for (int k = 0; k < N; ++k) {
            // largest |diagonal| entry of the trailing block, first one wins ties (maxCoeff visitor: value > res)
            int big = k;
            double best = std::fabs(m[k][k]);
            for (int i = k + 1; i < N; ++i) {
                double v = std::fabs(m[i][i]);
                if (v > best) { best = v; big = i; }
            }
            tr[k] = big;
            if (k != big) {
                // symmetric row/column swap touching the lower triangle only
                int s = N - big - 1;
                for (int j = 0; j < k; ++j) { double t = m[k][j]; m[k][j] = m[big][j]; m[big][j] = t; }
                for (int i = 0; i < s; ++i) {
                    double t = m[big + 1 + i][k]; m[big + 1 + i][k] = m[big + 1 + i][big]; m[big + 1 + i][big] = t;
                }
                { double t = m[k][k]; m[k][k] = m[big][big]; m[big][big] = t; }
                for (int i = k + 1; i < big; ++i) { double t = m[i][k]; m[i][k] = m[big][i]; m[big][i] = t; }
            }
            int rs = N - k - 1;
            if (k > 0) {
                // temp = D(0..k) .* A10^T ; A11 -= A10 * temp ; A21 -= A20 * temp
                double temp[N];
                for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
                double acc = 0;
                for (int j = 0; j < k; ++j) acc = (j == 0) ? m[k][j] * temp[j] : acc + m[k][j] * temp[j];
                m[k][k] = m[k][k] - acc;
                for (int i = 0; i < rs; ++i) {
                    double a2 = 0;
                    for (int j = 0; j < k; ++j) a2 = (j == 0) ? m[k + 1 + i][j] * temp[j] : a2 + m[k + 1 + i][j] * temp[j];
                    m[k + 1 + i][k] = m[k + 1 + i][k] - a2;
                }
            }
            double akk = m[k][k];
            bool pivot_is_valid = std::fabs(akk) > 0.0;
            if (k == 0 && !pivot_is_valid) {
                // "The entire diagonal is zero, there is nothing more to do except filling the transpositions"
                for (int j = 0; j < N; ++j) tr[j] = j;
                return;
            }
            if (rs > 0 && pivot_is_valid)
                for (int i = 0; i < rs; ++i) m[k + 1 + i][k] = m[k + 1 + i][k] / akk;
            if (!pivot_is_valid) found_zero_pivot = true;
        }
    }
    double m[N][N];
    int tr[N];
};

// ---- declared-only types for legoslam::triangulation (algorithm.h:11-34); never instantiated at run time ----
struct Loose {
    Loose();
    Loose(unsigned long, int);
    explicit Loose(unsigned long);
    void setZero();
    Loose row(int) const;
    Loose col(int) const;
    template <int R, int C> Loose &block(unsigned long, int);
    Loose &operator=(const Loose &);
    Loose bdcSvd(int) const;
    Loose matrixV() const;
    Loose singularValues() const;
    double operator()(int, int) const;
    double operator[](int) const;
    template <int K> Loose head() const;
};
Loose operator*(double, const Loose &);
Loose operator-(const Loose &, const Loose &);
Loose operator/(const Loose &, double);
struct LooseVec3 {
    double v[3];
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    LooseVec3 &operator=(const Loose &);
};
struct LooseSE3 {
    Loose matrix3x4() const;
};

}  // namespace ref_stub

enum { ComputeThinU = 0x08, ComputeThinV = 0x20 };

class Matrix2d {
public:
    Matrix2d() { m[0][0] = m[0][1] = m[1][0] = m[1][1] = 0; }
    static Matrix2d Zero() { return Matrix2d(); }
    double &operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
    Matrix2d &operator+=(const Matrix2d &o) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) m[i][j] = m[i][j] + o.m[i][j];
        return *this;
    }
    ref_stub::LDLT<2> ldlt() const { return ref_stub::LDLT<2>(m); }
    double m[2][2];
};

// outer product J * J^T, coefficient (i, j) = a[i] * b[j]
static inline Matrix2d operator*(const Vector2d &a, const RowVector2d &b) {
    Matrix2d r;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) r(i, j) = a[i] * b.v[j];
    return r;
}

}  // namespace Eigen
#endif
