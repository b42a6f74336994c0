// hostmath.h -- Eigen-free host/device math for the LSD-SLAM hot path: SE3 (unit quaternion + translation)
// in float and double, Sim3 helpers, 3x3 inverse, pivoted 6x6 LDL^T.
//
// Mirrors what the reference gets from Sophus/Eigen:
//   SE3Group::exp / operator* / inverse      thirdparty/Sophus/sophus/se3.hpp:406-428, 239-259, 167-172
//   SO3Group::expAndTheta / normalize        thirdparty/Sophus/sophus/so3.hpp:342-369, 196-202
//   A.ldlt().solve(b)                        Tracking/SE3Tracker.cpp:359 (Eigen LDLT, diagonal pivoting)
//   K.inverse()                              Tracking/SE3Tracker.cpp:60, DataStructures/Frame.cpp:409,453
// Usable from host and device code (the device-resident LM loop solves and updates the pose on the GPU).
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define LSD_HD __host__ __device__ __forceinline__
#else
#define LSD_HD inline
#endif

namespace lsd {

template <typename T> struct MathFn;
template <> struct MathFn<float> {
    static LSD_HD float sqrt_(float x) { return sqrtf(x); }
    static LSD_HD float sin_(float x) { return sinf(x); }
    static LSD_HD float cos_(float x) { return cosf(x); }
    static LSD_HD float eps() { return 1e-5f; }          // SophusConstants<float>::epsilon, sophus.hpp:52-56
};
template <> struct MathFn<double> {
    static LSD_HD double sqrt_(double x) { return sqrt(x); }
    static LSD_HD double sin_(double x) { return sin(x); }
    static LSD_HD double cos_(double x) { return cos(x); }
    static LSD_HD double eps() { return 1e-10; }          // sophus.hpp:43-47
};

// q = (x, y, z, w) like Eigen::Quaternion::coeffs()
template <typename T> struct SE3 {
    T q[4];
    T t[3];
    LSD_HD SE3() { q[0] = q[1] = q[2] = 0; q[3] = 1; t[0] = t[1] = t[2] = 0; }
};

template <typename T> LSD_HD void quatNormalize(T q[4])
{
    // coeffs().norm(): Eigen's unrolled reduction tree over 4 coefficients, (x^2 + y^2) + (z^2 + w^2)
    T len = MathFn<T>::sqrt_((q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]));
    q[0] /= len; q[1] /= len; q[2] /= len; q[3] /= len;
}

template <typename T> LSD_HD void quatMul(const T a[4], const T b[4], T o[4])
{
    T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    T y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    T z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

// v' = q v q*   (Eigen's _transformVector: two cross products)
template <typename T> LSD_HD void quatRotate(const T q[4], const T v[3], T o[3])
{
    T ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    T cx = q[1] * uz - q[2] * uy, cy = q[2] * ux - q[0] * uz, cz = q[0] * uy - q[1] * ux;
    o[0] = v[0] + q[3] * ux + cx;
    o[1] = v[1] + q[3] * uy + cy;
    o[2] = v[2] + q[3] * uz + cz;
}

template <typename T> LSD_HD void quatToMatrix(const T q[4], T R[9])
{
    T tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

template <typename T> LSD_HD SE3<T> se3Mul(const SE3<T>& a, const SE3<T>& b)
{
    SE3<T> r;
    T rt[3];
    quatRotate(a.q, b.t, rt);
    quatMul(a.q, b.q, r.q);
    quatNormalize(r.q);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    return r;
}

template <typename T> LSD_HD SE3<T> se3Inverse(const SE3<T>& a)
{
    SE3<T> r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    quatNormalize(r.q);
    T nt[3] = { a.t[0] * (T)-1, a.t[1] * (T)-1, a.t[2] * (T)-1 };
    quatRotate(r.q, nt, r.t);
    return r;
}

// exp of a twist a = [upsilon | omega]
template <typename T> LSD_HD SE3<T> se3Exp(const T a[6])
{
    typedef MathFn<T> M;
    const T* om = a + 3;
    T theta_sq = om[0] * om[0] + (om[1] * om[1] + om[2] * om[2]);   // omega.squaredNorm(): 1 + 2 reduction tree
    T theta = M::sqrt_(theta_sq);
    T half_theta = (T)0.5 * theta;
    T imag, real;
    if (theta < M::eps()) {
        T theta_po4 = theta_sq * theta_sq;
        imag = (T)0.5 - (T)(1.0 / 48.0) * theta_sq + (T)(1.0 / 3840.0) * theta_po4;
        real = (T)1 - (T)0.5 * theta_sq + (T)(1.0 / 384.0) * theta_po4;
    } else {
        imag = M::sin_(half_theta) / theta;
        real = M::cos_(half_theta);
    }
    SE3<T> r;
    r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
    quatNormalize(r.q);
    T Om[9] = { 0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0 };
    T Om2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Om2[i * 3 + j] = (Om[i * 3 + 0] * Om[0 * 3 + j] + Om[i * 3 + 1] * Om[1 * 3 + j]) + Om[i * 3 + 2] * Om[2 * 3 + j];
    T V[9];
    if (theta < M::eps()) {
        quatToMatrix(r.q, V);
    } else {
        T tsq = theta * theta;                                      // se3.hpp:419 recomputes theta_sq from theta
        T c1 = ((T)1 - M::cos_(theta)) / tsq;
        T c2 = (theta - M::sin_(theta)) / (tsq * theta);
        for (int i = 0; i < 9; i++) V[i] = (((i % 4) == 0 ? (T)1 : (T)0) + c1 * Om[i]) + c2 * Om2[i];
    }
    for (int i = 0; i < 3; i++) r.t[i] = (V[i * 3 + 0] * a[0] + V[i * 3 + 1] * a[1]) + V[i * 3 + 2] * a[2];
    return r;
}

// SE3::cast<>() : component-wise cast followed by the normalising constructor
template <typename To, typename From> LSD_HD SE3<To> se3Cast(const SE3<From>& a)
{
    SE3<To> r;
    for (int i = 0; i < 4; i++) r.q[i] = (To)a.q[i];
    quatNormalize(r.q);
    for (int i = 0; i < 3; i++) r.t[i] = (To)a.t[i];
    return r;
}

// cofactor inverse of a row-major 3x3 (what Eigen does for fixed size 3)
LSD_HD void mat3Inverse(const float m[9], float r[9])
{
    float c00 = m[4] * m[8] - m[5] * m[7];
    float c10 = m[7] * m[2] - m[8] * m[1];   // cofactor (1,0): rows 2,0 / cols 1,2
    float c20 = m[1] * m[5] - m[2] * m[4];
    float det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
    float inv = 1.0f / det;
    r[0] = c00 * inv; r[1] = c10 * inv; r[2] = c20 * inv;
    r[3] = (m[5] * m[6] - m[3] * m[8]) * inv;
    r[4] = (m[8] * m[0] - m[6] * m[2]) * inv;
    r[5] = (m[2] * m[3] - m[0] * m[5]) * inv;
    r[6] = (m[3] * m[7] - m[4] * m[6]) * inv;
    r[7] = (m[6] * m[1] - m[7] * m[0]) * inv;
    r[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
}

// Eigen's unrolled reduction tree (Redux.h): sum(first n/2) + sum(rest), written out for n <= 8 (no recursion: the
// function must inline into device code, otherwise the caller's matrices are forced into local memory)
LSD_HD float treeSumF(const float* v, int n)
{
    switch (n) {
    case 1: return v[0];
    case 2: return v[0] + v[1];
    case 3: return v[0] + (v[1] + v[2]);
    case 4: return (v[0] + v[1]) + (v[2] + v[3]);
    case 5: return (v[0] + v[1]) + (v[2] + (v[3] + v[4]));
    case 6: return (v[0] + (v[1] + v[2])) + (v[3] + (v[4] + v[5]));
    case 7: return (v[0] + (v[1] + v[2])) + ((v[3] + v[4]) + (v[5] + v[6]));
    default: return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
}

// x = A^-1 b for a symmetric NxN via LDL^T with largest-diagonal pivoting (float): Eigen's LDLT (unblocked lower
// factorisation; solve = P, unrolled unit-lower solve with tree sums, pseudo-inverse of D, unrolled upper solve, P^T)
template <int N> LSD_HD void ldltSolve(const float* Ain, const float* bin, float* x)
{
    float A[N][N];
    int tr[N];
    float cutoff = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) A[i][j] = Ain[i * N + j];
    for (int k = 0; k < N; k++) {
        int p = k;
        float big = fabsf(A[k][k]);
        for (int i = k + 1; i < N; i++)
            if (fabsf(A[i][i]) > big) { big = fabsf(A[i][i]); p = i; }
        if (k == 0) cutoff = fabsf(1.1920929e-07f * big);
        tr[k] = p;
        if (big < cutoff) { for (int i = k; i < N; i++) tr[i] = i; break; }
        if (p != k) {
            for (int j = 0; j < k; j++) { float t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
            for (int i = p + 1; i < N; i++) { float t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
            for (int i = k + 1; i < p; i++) { float t = A[i][k]; A[i][k] = A[p][i]; A[p][i] = t; }
            float t = A[k][k]; A[k][k] = A[p][p]; A[p][p] = t;
        }
        if (k > 0) {
            float temp[N];
            for (int j = 0; j < k; j++) temp[j] = A[j][j] * A[k][j];
            float s = 0;
            for (int j = 0; j < k; j++) s += A[k][j] * temp[j];
            A[k][k] -= s;
            for (int i = k + 1; i < N; i++) {
                float u = 0;
                for (int j = 0; j < k; j++) u += A[i][j] * temp[j];
                A[i][k] -= u;
            }
        }
        float d = A[k][k];
        if (fabsf(d) > cutoff)
            for (int i = k + 1; i < N; i++) A[i][k] /= d;
    }
    float y[N];
    for (int i = 0; i < N; i++) y[i] = bin[i];
    for (int k = 0; k < N; k++)
        if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 1; i < N; i++) {
        float pr[N];
        for (int j = 0; j < i; j++) pr[j] = A[i][j] * y[j];
        y[i] -= treeSumF(pr, i);
    }
    float dmax = 0;
    for (int i = 0; i < N; i++) if (fabsf(A[i][i]) > dmax) dmax = fabsf(A[i][i]);
    float tol = dmax * 1.1920929e-07f;
    if (tol < 1.0f / 3.40282347e+38f) tol = 1.0f / 3.40282347e+38f;
    for (int i = 0; i < N; i++) { float d = A[i][i]; y[i] = (fabsf(d) > tol) ? y[i] / d : 0.0f; }
    for (int i = N - 2; i >= 0; i--) {
        float pr[N];
        const int n = N - 1 - i;
        for (int j = 0; j < n; j++) pr[j] = A[i + 1 + j][i] * y[i + 1 + j];
        y[i] -= treeSumF(pr, n);
    }
    for (int k = N - 1; k >= 0; k--)
        if (tr[k] != k) { float t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < N; i++) x[i] = y[i];
}
LSD_HD void ldlt6Solve(const float Ain[36], const float bin[6], float x[6]) { ldltSolve<6>(Ain, bin, x); }

// ---- Sim3 as Sophus stores it: NON-unit quaternion (|q| = scale) + translation, double -------------------------
// thirdparty/Sophus/sophus/sim3.hpp:159-173 (inverse), :257-260 (operator*=), :418-428 (exp), :609-648 (calcW);
// rxso3.hpp:194-200 (inverse), :221-228 (matrix), :262-269 (operator* on a point), :299-305 (rotationMatrix), :416-425 (exp)
struct Sim3 {
    double q[4];
    double t[3];
    LSD_HD Sim3() { q[0] = q[1] = q[2] = 0; q[3] = 1; t[0] = t[1] = t[2] = 0; }
};
LSD_HD double sim3Scale(const Sim3& a) { return sqrt(a.q[0] * a.q[0] + a.q[1] * a.q[1] + a.q[2] * a.q[2] + a.q[3] * a.q[3]); }
// qts[8] = unit quaternion (x,y,z,w), translation, scale  (the ABI's Sim3 layout, like thisToParent_qts)
LSD_HD Sim3 sim3FromQts(const double a[8])
{
    Sim3 r;
    for (int i = 0; i < 4; i++) r.q[i] = a[i] * a[7];
    for (int i = 0; i < 3; i++) r.t[i] = a[4 + i];
    return r;
}
LSD_HD void sim3ToQts(const Sim3& a, double o[8])
{
    const double s = sim3Scale(a);
    for (int i = 0; i < 4; i++) o[i] = a.q[i] / s;
    for (int i = 0; i < 3; i++) o[4 + i] = a.t[i];
    o[7] = s;
}
LSD_HD void rxso3Apply(const double q[4], const double p[3], double o[3])
{
    const double scale = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double nq[4] = { q[0] / scale, q[1] / scale, q[2] / scale, q[3] / scale };
    double r[3];
    quatRotate(nq, p, r);
    o[0] = scale * r[0]; o[1] = scale * r[1]; o[2] = scale * r[2];
}
LSD_HD Sim3 sim3Mul(const Sim3& a, const Sim3& b)
{
    Sim3 r;
    double rt[3];
    rxso3Apply(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    quatMul(a.q, b.q, r.q);
    return r;
}
LSD_HD Sim3 sim3Inverse(const Sim3& a)
{
    const double n2 = a.q[0] * a.q[0] + a.q[1] * a.q[1] + a.q[2] * a.q[2] + a.q[3] * a.q[3];
    Sim3 r;
    r.q[0] = -a.q[0] / n2; r.q[1] = -a.q[1] / n2; r.q[2] = -a.q[2] / n2; r.q[3] = a.q[3] / n2;
    const double nt[3] = { a.t[0] * -1.0, a.t[1] * -1.0, a.t[2] * -1.0 };
    rxso3Apply(r.q, nt, r.t);
    return r;
}
// exp of a = [upsilon | omega | sigma]
LSD_HD Sim3 sim3Exp(const double a[7])
{
    const double* ups = a;
    const double* om = a + 3;
    const double sigma = a[6];
    const double scale = exp(sigma);
    const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double theta = sqrt(theta_sq), half_theta = 0.5 * theta;
    double imag, real;
    if (theta < 1e-10) {
        const double theta_po4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        imag = sin(half_theta) / theta;
        real = cos(half_theta);
    }
    Sim3 r;
    r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
    quatNormalize(r.q);
    for (int i = 0; i < 4; i++) r.q[i] *= scale;
    const double rs = sim3Scale(r);
    const double Om[9] = { 0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0 };
    double Om2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Om2[i * 3 + j] = (Om[i * 3 + 0] * Om[0 * 3 + j] + Om[i * 3 + 1] * Om[1 * 3 + j]) + Om[i * 3 + 2] * Om[2 * 3 + j];
    double A, B, C;
    if (fabs(sigma) < 1e-10) {
        C = 1.0;
        if (fabs(theta) < 1e-10) { A = 0.5; B = 1.0 / 6.0; }
        else { A = (1.0 - cos(theta)) / theta_sq; B = (theta - sin(theta)) / (theta_sq * theta); }
    } else {
        C = (rs - 1.0) / sigma;
        if (fabs(theta) < 1e-10) {
            const double sigma_sq = sigma * sigma;
            A = ((sigma - 1.0) * rs + 1.0) / sigma_sq;
            B = ((0.5 * sigma * sigma - sigma + 1.0) * rs) / (sigma_sq * sigma);
        } else {
            const double sa = rs * sin(theta), sb = rs * cos(theta), sc = theta_sq + sigma * sigma;
            A = (sa * sigma + (1.0 - sb) * theta) / (theta * sc);
            B = (C - ((sb - 1.0) * sigma + sa * theta) / (sc)) * 1.0 / (theta_sq);
        }
    }
    for (int i = 0; i < 3; i++) {
        double w[3];
        for (int j = 0; j < 3; j++) w[j] = (A * Om[i * 3 + j] + B * Om2[i * 3 + j]) + C * (i == j ? 1.0 : 0.0);
        r.t[i] = (w[0] * ups[0] + w[1] * ups[1]) + w[2] * ups[2];
    }
    return r;
}
// The per-pose constants of Sim3Tracker::calcSim3Buffers (Sim3Tracker.cpp:447-460): rxso3().matrix(), translation and
// the roll of the unscaled rotation about the optical axis (Eigen Quaternion::setFromTwoVectors + toRotationMatrix),
// all cast to float as the reference does.  The 180-degree SVD branch of setFromTwoVectors is not restated (NaN).
LSD_HD void sim3PoseConstants(const Sim3& a, float rotMat[9], float transVec[3], float roll[4])
{
    const double scale = sim3Scale(a);
    const double nq[4] = { a.q[0] / scale, a.q[1] / scale, a.q[2] / scale, a.q[3] / scale };
    double R[9];
    quatToMatrix(nq, R);
    float Ru[9];
    for (int i = 0; i < 9; i++) { rotMat[i] = (float)(scale * R[i]); Ru[i] = (float)R[i]; }
    for (int i = 0; i < 3; i++) transVec[i] = (float)a.t[i];
    const float rf[3] = { (Ru[0] * 0.f + Ru[1] * 0.f) + Ru[2] * -1.f, (Ru[3] * 0.f + Ru[4] * 0.f) + Ru[5] * -1.f, (Ru[6] * 0.f + Ru[7] * 0.f) + Ru[8] * -1.f };
    const float n = sqrtf((rf[0] * rf[0] + rf[1] * rf[1]) + rf[2] * rf[2]);
    const float v0[3] = { rf[0] / n, rf[1] / n, rf[2] / n }, v1[3] = { 0.f, 0.f, -1.f };
    const float c = (v1[0] * v0[0] + v1[1] * v0[1]) + v1[2] * v0[2];
    if (c < -1.0f + 1e-5f) { roll[0] = roll[1] = roll[2] = roll[3] = nanf(""); return; }
    const float ax[3] = { v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0] };
    const float s = sqrtf((1.f + c) * 2.f), invs = 1.f / s;
    const float q[4] = { ax[0] * invs, ax[1] * invs, ax[2] * invs, s * 0.5f };
    float Rb[9];
    quatToMatrix(q, Rb);
    roll[0] = (Rb[0] * Ru[0] + Rb[1] * Ru[3]) + Rb[2] * Ru[6];
    roll[1] = (Rb[0] * Ru[1] + Rb[1] * Ru[4]) + Rb[2] * Ru[7];
    roll[2] = (Rb[3] * Ru[0] + Rb[4] * Ru[3]) + Rb[5] * Ru[6];
    roll[3] = (Rb[3] * Ru[1] + Rb[4] * Ru[4]) + Rb[5] * Ru[7];
}

}  // namespace lsd
