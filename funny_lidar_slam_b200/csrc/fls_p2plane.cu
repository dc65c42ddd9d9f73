// fls_p2plane.cu — K1: iVox 5-NN + plane fit + point-to-plane Jacobian/residual + block reduction.
//
// One thread per source point fuses what LoamPointToPlaneIVOX::PlanerMatch and ::SumCoefficient do
// (include/registration/loam_point_to_plane_ivox.h:256-340 upstream): transform the point with the current
// pose, probe the 19-voxel stencil of the hash grid, keep the 5 nearest map points, least-squares plane
// through them (column-pivoted Householder QR, fp64), validity / near-point gates, J (6) and |d|, and the
// accumulation of the 21+6+2 Gauss-Newton sums — reduced per block into one row of the partial matrix.
//
// State that survives across iterations of one Match [quirk 1, SURVEY.md §7]: upstream resets the valid flags
// once per Match and sums every flagged point, so a point that was valid earlier but is rejected now keeps
// contributing its stale H_i, g_i.  The kernel therefore writes a 32-byte record {J[6], |d|} per valid point
// (SoA: two float4 arrays) plus a flag byte, and re-reads it only on the stale path.
#include "fls_gn.cuh"
#include "fls_ivox.cuh"
#include "fls_kernels.h"

namespace fls {
namespace {

// Householder step on column K (rows K..4) of the 5x3 system, applied to the trailing columns and the rhs.
template <int K>
__device__ __forceinline__ void hh_step(double (&A)[5][3], double (&b)[5]) {
    const double alpha = A[K][K];
    double tail = 0;
#pragma unroll
    for (int i = K + 1; i < 5; ++i) tail += A[i][K] * A[i][K];
    if (tail == 0.0) return;  // tau = 0, beta = alpha: nothing to apply
    double beta = sqrt(alpha * alpha + tail);
    if (alpha >= 0) beta = -beta;
    const double inv = 1.0 / (alpha - beta);
    const double tau = (beta - alpha) / beta;
    double v[5];
    v[K] = 1.0;
#pragma unroll
    for (int i = K + 1; i < 5; ++i) v[i] = A[i][K] * inv;
    A[K][K] = beta;
#pragma unroll
    for (int j = K + 1; j < 3; ++j) {
        double s = 0;
#pragma unroll
        for (int i = K; i < 5; ++i) s += v[i] * A[i][j];
        s *= tau;
#pragma unroll
        for (int i = K; i < 5; ++i) A[i][j] -= s * v[i];
    }
    double s = 0;
#pragma unroll
    for (int i = K; i < 5; ++i) s += v[i] * b[i];
    s *= tau;
#pragma unroll
    for (int i = K; i < 5; ++i) b[i] -= s * v[i];
}

template <int K>
__device__ __forceinline__ double colnorm2(const double (&A)[5][3], int j) {
    double s = 0;
#pragma unroll
    for (int i = K; i < 5; ++i) s += (j == 0 ? A[i][0] : (j == 1 ? A[i][1] : A[i][2])) * (j == 0 ? A[i][0] : (j == 1 ? A[i][1] : A[i][2]));
    return s;
}
__device__ __forceinline__ void swap_cols(double (&A)[5][3], int a, int bcol) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double t = A[i][a];
        A[i][a] = A[i][bcol];
        A[i][bcol] = t;
    }
}

// min || A c + 1 ||  — Eigen colPivHouseholderQr().solve(b) with b = -1 (loam_point_to_plane_ivox.h:275-283 upstream)
__device__ __forceinline__ void plane_lstsq(double (&A)[5][3], double (&c)[3]) {
    double b[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
    int p0 = 0, p1 = 1, p2 = 2;
    int rank = 3;
    double n0 = colnorm2<0>(A, 0), n1 = colnorm2<0>(A, 1), n2 = colnorm2<0>(A, 2);
    const double maxcn = fmax(n0, fmax(n1, n2));
    const double th = 2.220446049250313e-16 * sqrt(maxcn) / 5.0;
    const double thr = th * th;
    // k = 0
    {
        int piv = 0;
        double best = n0;
        if (n1 > best) { best = n1; piv = 1; }
        if (n2 > best) { best = n2; piv = 2; }
        if (best < thr || best == 0.0) {
            rank = 0;
        } else {
            if (piv == 1) { swap_cols(A, 0, 1); int t = p0; p0 = p1; p1 = t; }
            if (piv == 2) { swap_cols(A, 0, 2); int t = p0; p0 = p2; p2 = t; }
            hh_step<0>(A, b);
        }
    }
    if (rank == 3) {  // k = 1
        n1 = colnorm2<1>(A, 1);
        n2 = colnorm2<1>(A, 2);
        double best = n1;
        int piv = 1;
        if (n2 > best) { best = n2; piv = 2; }
        if (best < thr || best == 0.0) {
            rank = 1;
        } else {
            if (piv == 2) { swap_cols(A, 1, 2); int t = p1; p1 = p2; p2 = t; }
            hh_step<1>(A, b);
        }
    }
    if (rank == 3) {  // k = 2
        n2 = colnorm2<2>(A, 2);
        if (n2 < thr || n2 == 0.0) rank = 2;
        else hh_step<2>(A, b);
    }
    const double y2 = (rank > 2) ? b[2] / A[2][2] : 0.0;
    const double y1 = (rank > 1) ? (b[1] - A[1][2] * y2) / A[1][1] : 0.0;
    const double y0 = (rank > 0) ? (b[0] - A[0][1] * y1 - A[0][2] * y2) / A[0][0] : 0.0;
    c[0] = (p0 == 0) ? y0 : ((p1 == 0) ? y1 : y2);
    c[1] = (p0 == 1) ? y0 : ((p1 == 1) ? y1 : y2);
    c[2] = (p0 == 2) ? y0 : ((p1 == 2) ? y1 : y2);
}

// Geometry of one source point against the map: returns true when the point produces a valid residual.
__device__ __forceinline__ bool p2plane_point(const IvoxView& map, const float4 sp, const double* __restrict__ pose /*R[9], t[3]*/,
                                              double plane_thres, double (&J)[6], double& ad, unsigned& n_cand, unsigned& n_hits) {
    const double px = sp.x, py = sp.y, pz = sp.z;
    const float qx = xform_row_d(pose[0], pose[1], pose[2], pose[9], px, py, pz);
    const float qy = xform_row_d(pose[3], pose[4], pose[5], pose[10], px, py, pz);
    const float qz = xform_row_d(pose[6], pose[7], pose[8], pose[11], px, py, pz);
    Knn5 nn;
    ivox_knn5(map, qx, qy, qz, nn, n_cand, n_hits);
    if (nn.j4 == 0xffffffffu) return false;  // fewer than 5 neighbours (:271-273)
    double A[5][3];
    {
        const float4 a0 = __ldg(map.pts + nn.j0), a1 = __ldg(map.pts + nn.j1), a2 = __ldg(map.pts + nn.j2), a3 = __ldg(map.pts + nn.j3),
                     a4 = __ldg(map.pts + nn.j4);
        A[0][0] = a0.x; A[0][1] = a0.y; A[0][2] = a0.z;
        A[1][0] = a1.x; A[1][1] = a1.y; A[1][2] = a1.z;
        A[2][0] = a2.x; A[2][1] = a2.y; A[2][2] = a2.z;
        A[3][0] = a3.x; A[3][1] = a3.y; A[3][2] = a3.z;
        A[4][0] = a4.x; A[4][1] = a4.y; A[4][2] = a4.z;
    }
    const double a0x = A[0][0], a0y = A[0][1], a0z = A[0][2];
    double Aq[5][3];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Aq[i][j] = A[i][j];
    double c[3];
    plane_lstsq(Aq, c);
    const double cn = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    bool valid = true;
#pragma unroll
    for (int j = 0; j < 5; ++j)
        if (fabs(A[j][0] * c[0] + A[j][1] * c[1] + A[j][2] * c[2] + 1.0) / cn > plane_thres) valid = false;  // :286-293
    if (!valid) return false;
    const double nx = c[0] / cn, ny = c[1] / cn, nz = c[2] / cn;
    const double d = ((double)qx - a0x) * nx + ((double)qy - a0y) * ny + ((double)qz - a0z) * nz;  // :306 from the nearest neighbour
    if (sqrt(px * px + py * py + pz * pz) < 81.0 * d * d) return false;                              // :309 body-frame norm
    const double s = d > 0 ? 1.0 : -1.0;
    const double rx = pose[0] * px + pose[1] * py + pose[2] * pz;
    const double ry = pose[3] * px + pose[4] * py + pose[5] * pz;
    const double rz = pose[6] * px + pose[7] * py + pose[8] * pz;
    J[0] = s * (ry * nz - rz * ny);  // (R p) x n  == -hat(R p)^T n  (:315)
    J[1] = s * (rz * nx - rx * nz);
    J[2] = s * (rx * ny - ry * nx);
    J[3] = s * nx;
    J[4] = s * ny;
    J[5] = s * nz;
    ad = fabs(d);
    return true;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) p2plane_iter_kernel(P2PlaneArgs a) {
    __shared__ double s_pose[12];
    if (a.state->done) return;  // uniform: loop already finished on the device
    if (threadIdx.x < 9) s_pose[threadIdx.x] = a.state->R[threadIdx.x];
    else if (threadIdx.x < 12) s_pose[threadIdx.x] = a.state->t[threadIdx.x - 9];
    __syncthreads();

    double acc[kNumAcc];
#pragma unroll
    for (int k = 0; k < kNumAcc; ++k) acc[k] = 0.0;

    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i < a.n) {
        const float4 sp = a.src[i];
        double J[6], ad = 0;
        unsigned n_cand, n_hits;
        bool use = p2plane_point(a.map, sp, s_pose, a.plane_thres, J, ad, n_cand, n_hits);
        acc[kAccCand] = (double)n_cand;
        acc[kAccHits] = (double)n_hits;
        if (use) {
            a.rec0[i] = make_float4((float)J[0], (float)J[1], (float)J[2], (float)J[3]);
            a.rec1[i] = make_float4((float)J[4], (float)J[5], (float)ad, 1.0f);
            a.flags[i] = 1;
        } else if (a.flags[i]) {  // stale contribution [quirk 1]
            const float4 r0 = a.rec0[i], r1 = a.rec1[i];
            J[0] = r0.x; J[1] = r0.y; J[2] = r0.z; J[3] = r0.w; J[4] = r1.x; J[5] = r1.y;
            ad = r1.z;
            use = true;
        }
        if (use) {
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = r; c < 6; ++c) acc[k++] += J[r] * J[c];
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[21 + r] += -J[r] * ad;
            acc[kAccValid] += 1.0;
            acc[kAccRes] += ad;
        }
    }
    block_reduce_store<BLOCK>(acc, a.partials + (size_t)blockIdx.x * kAccStride);
}

__global__ void ivox_knn_test_kernel(IvoxView map, const float4* __restrict__ q, int n, float4* __restrict__ out, int* __restrict__ found) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = q[i];
    Knn5 nn;
    unsigned nc, nh;
    ivox_knn5(map, p.x, p.y, p.z, nn, nc, nh);
    const unsigned js[5] = {nn.j0, nn.j1, nn.j2, nn.j3, nn.j4};
    int f = 0;
    for (int k = 0; k < 5; ++k) {
        if (js[k] != 0xffffffffu) {
            out[(size_t)i * 5 + k] = map.pts[js[k]];
            ++f;
        } else {
            out[(size_t)i * 5 + k] = make_float4(0, 0, 0, 0);
        }
    }
    found[i] = f;
}

}  // namespace

int p2plane_grid(int n) { return (n + kP2PlaneBlock - 1) / kP2PlaneBlock; }

void launch_p2plane_iter(const P2PlaneArgs& a, cudaStream_t st) {
    if (a.n <= 0) return;
    p2plane_iter_kernel<kP2PlaneBlock><<<p2plane_grid(a.n), kP2PlaneBlock, 0, st>>>(a);
}

void launch_ivox_knn_test(const IvoxView& map, const float4* d_q, int n, float4* d_out, int* d_found, cudaStream_t st) {
    if (n <= 0) return;
    ivox_knn_test_kernel<<<(n + 127) / 128, 128, 0, st>>>(map, d_q, n, d_out, d_found);
}

}  // namespace fls
