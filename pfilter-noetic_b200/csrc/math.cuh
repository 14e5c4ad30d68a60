// Register-resident fp64 helpers for the association and pose-solve kernels: quaternion rotation, symmetric 3x3
// eigen-solver (line fit), 5x3 column-pivoted Householder least squares (plane fit), SE(3) exponential / Plus,
// 6x6 Cholesky.  References are to /root/reference/src/odomEstimationClass.cpp and src/lidarOptimization.cpp.
// All code is written for nvcc -fmad=false (no contraction).
#pragma once
#include "common.cuh"

namespace pf {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double norm3(D3 a) { return sqrt(dot3(a, a)); }

// Eigen's quaternion * vector: v + w * uv + qv x uv with uv = 2 (qv x v); q = [x y z w]
__device__ __forceinline__ D3 quat_rotate(const double* q, D3 v) {
    D3 qv = d3(q[0], q[1], q[2]);
    D3 uv = cross3(qv, v);
    uv = uv + uv;
    return v + (q[3] * uv) + cross3(qv, uv);
}
// pointAssociateToMap (:162-168): p_w = q * p + t in double
__device__ __forceinline__ D3 pose_apply(const double* pose, D3 p) { return quat_rotate(pose, p) + d3(pose[4], pose[5], pose[6]); }

// ---------------------------------------------------------------------------------------------------------
// symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations (fp64).  Returns the eigenvalues in ascending order
// in w[] and the unit eigenvector of the largest one in vmax (what :321-331 consumes).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq, double& vp0, double& vq0,
                                              double& vp1, double& vq1, double& vp2, double& vq2) {
    if (apq == 0.0) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    app = app - t * apq;
    aqq = aqq + t * apq;
    apq = 0.0;
    const double rp = arp, rq = arq;   // the remaining off-diagonal pair (row r != p, q)
    arp = c * rp - s * rq;
    arq = s * rp + c * rq;
    double a, b;
    a = vp0; b = vq0; vp0 = c * a - s * b; vq0 = s * a + c * b;
    a = vp1; b = vq1; vp1 = c * a - s * b; vq1 = s * a + c * b;
    a = vp2; b = vq2; vp2 = c * a - s * b; vq2 = s * a + c * b;
}

__device__ __forceinline__ void eig3_sym(double a00, double a01, double a02, double a11, double a12, double a22, double w[3], D3& vmax) {
    // V = I, columns v0 v1 v2 stored by rows: v{col}{row}
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;   // v{c}{r}: component r of column c
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        const double dg = a00 * a00 + a11 * a11 + a22 * a22;
        if (off == 0.0 || off <= 1e-40 * dg) break;
        // (p,q) = (0,1): other row r = 2 -> pair (a02, a12)
        jacobi_rotate(a00, a11, a01, a02, a12, v00, v10, v01, v11, v02, v12);
        // (0,2): r = 1 -> pair (a01, a12) where a12 = a21
        jacobi_rotate(a00, a22, a02, a01, a12, v00, v20, v01, v21, v02, v22);
        // (1,2): r = 0 -> pair (a01, a02)
        jacobi_rotate(a11, a22, a12, a01, a02, v10, v20, v11, v21, v12, v22);
    }
    // ascending order
    double e0 = a00, e1 = a11, e2 = a22;
    D3 c0 = d3(v00, v01, v02), c1 = d3(v10, v11, v12), c2 = d3(v20, v21, v22);
    if (e0 > e1) { double t = e0; e0 = e1; e1 = t; D3 c = c0; c0 = c1; c1 = c; }
    if (e1 > e2) { double t = e1; e1 = e2; e2 = t; D3 c = c1; c1 = c2; c2 = c; }
    if (e0 > e1) { double t = e0; e0 = e1; e1 = t; D3 c = c0; c0 = c1; c1 = c; }
    w[0] = e0; w[1] = e1; w[2] = e2;
    vmax = c2;
}

// ---------------------------------------------------------------------------------------------------------
// x = argmin |A x - b| for the 5x3 neighbour matrix, column-pivoted Householder QR
// (Eigen::ColPivHouseholderQR::solve semantics used at :461; rank-deficient columns give zero components).
// A is passed by columns: c0[5], c1[5], c2[5].
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void plane_lsq_5x3(double (&A)[3][5], double (&b)[5], double x[3]) {
    int perm[3] = {0, 1, 2};
    double nu[3], nd[3], tau[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) s += A[k][i] * A[k][i];
        nu[k] = nd[k] = sqrt(s);
    }
    const double eps = 2.220446049250313e-16;
    const double mx = fmax(nu[0], fmax(nu[1], nu[2]));
    const double thr = (mx * eps / 5.0) * (mx * eps / 5.0);
    const double downdate = 1.4901161193847656e-08;   // sqrt(eps)
    int nonzero = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int big = k;
#pragma unroll
        for (int j = k + 1; j < 3; ++j) if (j > k && nu[j] > nu[big]) big = j;
        if (nonzero == 3 && nu[big] * nu[big] < thr * (double)(5 - k)) nonzero = k;
        if (big != k) {
#pragma unroll
            for (int i = 0; i < 5; ++i) { double t = A[k][i]; A[k][i] = A[big][i]; A[big][i] = t; }
            double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
            int p = perm[k]; perm[k] = perm[big]; perm[big] = p;
        }
        double tail = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) if (i > k) tail += A[k][i] * A[k][i];
        const double c0 = A[k][k];
        double beta, tk;
        if (tail <= 2.2250738585072014e-308) {
            tk = 0; beta = c0;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i > k) A[k][i] = 0;
        } else {
            beta = sqrt(c0 * c0 + tail);
            if (c0 >= 0) beta = -beta;
            const double den = c0 - beta;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i > k) A[k][i] = A[k][i] / den;
            tk = (beta - c0) / beta;
        }
        tau[k] = tk;
        A[k][k] = beta;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j > k) {
                double s = A[j][k];
#pragma unroll
                for (int i = 0; i < 5; ++i) if (i > k) s += A[k][i] * A[j][i];
                s *= tk;
                A[j][k] -= s;
#pragma unroll
                for (int i = 0; i < 5; ++i) if (i > k) A[j][i] -= s * A[k][i];
                if (nu[j] != 0) {   // LAPACK working note 176 norm down-date
                    double temp = fabs(A[j][k]) / nu[j];
                    temp = (1.0 + temp) * (1.0 - temp);
                    temp = temp < 0 ? 0 : temp;
                    const double r2 = nu[j] / nd[j];
                    if (temp * r2 * r2 <= downdate) {
                        double s2 = 0;
#pragma unroll
                        for (int i = 0; i < 5; ++i) if (i > k) s2 += A[j][i] * A[j][i];
                        nd[j] = sqrt(s2);
                        nu[j] = nd[j];
                    } else {
                        nu[j] *= sqrt(temp);
                    }
                }
            }
        }
        // b <- H_k b
        if (k < nonzero) {
            double s = b[k];
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i > k) s += A[k][i] * b[i];
            s *= tk;
            b[k] -= s;
#pragma unroll
            for (int i = 0; i < 5; ++i) if (i > k) b[i] -= s * A[k][i];
        }
    }
    double y[3] = {0, 0, 0};
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        if (i < nonzero) {
            double s = b[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j > i && j < nonzero) s -= A[j][i] * y[j];
            y[i] = s / A[i][i];
        }
    }
    x[0] = x[1] = x[2] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < nonzero) {
            if (perm[i] == 0) x[0] = y[i];
            else if (perm[i] == 1) x[1] = y[i];
            else x[2] = y[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// SE(3): getTransformFromSe3 (src/lidarOptimization.cpp:106-143) and PoseSE3Parameterization::Plus (:80-95)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_mul(const double* a, const double* b, double* o) {   // [x y z w]
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

__device__ __forceinline__ void quat_to_mat(const double* q, double* R) {   // Eigen toRotationMatrix, row-major
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void mat_to_quat(const double* R, double* q) {   // Eigen Quaternion(Matrix3)
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}

__device__ __forceinline__ void se3_exp(const double* d, double* dq, D3& dt) {
    const D3 om = d3(d[0], d[1], d[2]), up = d3(d[3], d[4], d[5]);
    const double theta = norm3(om), half = 0.5 * theta;
    double sh, ch;
    sincos(half, &sh, &ch);
    const double real = ch;
    double imag;
    if (theta < 1e-10) {
        const double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = sh / theta;
    }
    dq[0] = imag * om.x; dq[1] = imag * om.y; dq[2] = imag * om.z; dq[3] = real;
    double J[9];
    if (theta < 1e-10) {
        quat_to_mat(dq, J);
    } else {
        const double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
        double st, ct;
        sincos(theta, &st, &ct);
        const double c1 = (1 - ct) / (theta * theta), c2 = (theta - st) / (theta * theta * theta);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double o2 = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
                J[3 * i + j] = ((i == j) ? 1.0 : 0.0) + c1 * O[3 * i + j] + c2 * o2;
            }
    }
    dt = d3(J[0] * up.x + J[1] * up.y + J[2] * up.z, J[3] * up.x + J[4] * up.y + J[5] * up.z, J[6] * up.x + J[7] * up.y + J[8] * up.z);
}

__device__ __forceinline__ void se3_plus(const double* x, const double* d, double* out) {
    double dq[4];
    D3 dt;
    se3_exp(d, dq, dt);
    quat_mul(dq, x, out);   // quater_plus = delta_q * quater
    const D3 tp = quat_rotate(dq, d3(x[4], x[5], x[6])) + dt;
    out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

// Solve (A) y = b for a symmetric positive definite 6x6 (upper triangle U21, row-major) by Cholesky. false if not SPD.
// Fully unrolled so that the factor lives in registers.
__device__ __forceinline__ bool chol6_solve(const double* U21, const double* b, double* y) {
    double L[6][6];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { L[j][i] = U21[k]; ++k; }
    }
    bool ok = true;
    double inv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double s = L[j][j];
#pragma unroll
        for (int p = 0; p < 6; ++p) if (p < j) s -= L[j][p] * L[j][p];
        if (!(s > 0.0)) ok = false;
        const double d = sqrt(s);
        L[j][j] = d;
        inv[j] = 1.0 / d;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i > j) {
                double t = L[i][j];
#pragma unroll
                for (int p = 0; p < 6; ++p) if (p < j) t -= L[i][p] * L[j][p];
                L[i][j] = t * inv[j];
            }
        }
    }
    if (!ok) return false;
    double z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
#pragma unroll
        for (int p = 0; p < 6; ++p) if (p < i) s -= L[i][p] * z[p];
        z[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = z[i];
#pragma unroll
        for (int p = 0; p < 6; ++p) if (p > i) s -= L[p][i] * y[p];
        y[i] = s * inv[i];
    }
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 6; ++i) fin = fin && isfinite(y[i]);
    return fin;
}

}  // namespace pf
