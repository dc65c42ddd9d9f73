// fls_eig.cuh — eigen-decomposition of a symmetric 3x3 (cyclic Jacobi), the stand-in for Eigen::JacobiSVD on the PSD
// matrices upstream feeds it (incremental_ndt.h:166 voxel covariance, loam_full_kdtree.h:244 corner covariance).
#pragma once
#include "fls_common.cuh"

namespace fls {
namespace {

// cyclic Jacobi on a symmetric 3x3; eigenvalues descending, V columns (stands for JacobiSVD of a PSD matrix, :166)
__device__ __noinline__ void sym_eig3_dev(const double* S, double* lam, double* V) {
    double A[9];
    for (int i = 0; i < 9; ++i) {
        A[i] = S[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = c * akp - s * akq;
                    A[k * 3 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = c * apk - s * aqk;
                    A[q * 3 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq;
                    V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    const double d[3] = {A[0], A[4], A[8]};
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (d[idx[b]] > d[idx[a]]) {
                const int t = idx[a];
                idx[a] = idx[b];
                idx[b] = t;
            }
    double Vs[9];
    for (int j = 0; j < 3; ++j) {
        lam[j] = d[idx[j]];
        for (int i = 0; i < 3; ++i) Vs[i * 3 + j] = V[i * 3 + idx[j]];
    }
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

}  // namespace
}  // namespace fls
