// fls_p2plane.cu — K1 (+ fused K6): the whole LoamPointToPlaneIVOX Gauss-Newton loop as ONE persistent
// cooperative kernel.
//
// Per source point and iteration it fuses what LoamPointToPlaneIVOX::PlanerMatch / ::SumCoefficient do
// (include/registration/loam_point_to_plane_ivox.h:256-340 upstream): transform with the current pose, bounded
// 5-NN in the iVox map, least-squares plane through the 5 neighbours (column-pivoted Householder QR, fp64),
// validity / near-point gates, J (6) and |d|, and the 21+6+2 Gauss-Newton sums; then, after a grid-wide barrier,
// block 0 reduces the per-block partials in a fixed order, solves the 6x6 system, updates the pose and applies the
// stop rule (:167-203), and a second barrier releases the next iteration.  No host round trip inside a Match.
//
// B200 mapping
//   * grid = #SMs x resident CTAs (cooperative launch), grid-stride over points: no tail wave, no relaunch gaps;
//   * queries are processed in Morton order of their voxel (sorted once per Match), so the lanes of a warp share
//     centre voxels: the table probe and the candidate stream are the same addresses -> L1 broadcast, no divergence;
//   * k-NN = 1 probe of the centre table + a streaming scan of that centre's contiguous stencil list (fls_ivox.cuh);
//   * the 29 sums are accumulated warp-transposed: each lane stages {J, |d|, flags} in shared memory and lane k then
//     owns sum k (32 FMAs on broadcast LDS) — one register pair of accumulator state instead of 62, no shuffles;
//   * state that survives across iterations [quirk 1, SURVEY.md §7]: upstream resets the valid flags once per Match
//     and sums every flagged point, so a point valid earlier but rejected now keeps contributing its stale H_i, g_i:
//     a 32-byte record {J[6], |d|} per valid point (two float4 arrays) + a flag byte, re-read only on the stale path.
#include <cooperative_groups.h>

#include <cub/cub.cuh>

#include "fls_gn.cuh"
#include "fls_ivox.cuh"
#include "fls_kernels.h"

namespace cg = cooperative_groups;

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

template <int K, int J>
__device__ __forceinline__ double colnorm2(const double (&A)[5][3]) {
    double s = 0;
#pragma unroll
    for (int i = K; i < 5; ++i) s += A[i][J] * A[i][J];
    return s;
}
template <int CA, int CB>
__device__ __forceinline__ void swap_cols(double (&A)[5][3]) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double t = A[i][CA];
        A[i][CA] = A[i][CB];
        A[i][CB] = t;
    }
}

// min || A c + 1 ||  — Eigen colPivHouseholderQr().solve(b) with b = -1 (loam_point_to_plane_ivox.h:275-283 upstream)
__device__ __forceinline__ void plane_lstsq(double (&A)[5][3], double (&c)[3]) {
    double b[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
    int p0 = 0, p1 = 1, p2 = 2;
    int rank = 3;
    double n0 = colnorm2<0, 0>(A), n1 = colnorm2<0, 1>(A), n2 = colnorm2<0, 2>(A);
    const double maxcn = fmax(n0, fmax(n1, n2));
    const double th = 2.220446049250313e-16 * sqrt(maxcn) / 5.0;
    const double thr = th * th;
    {  // k = 0
        int piv = 0;
        double best = n0;
        if (n1 > best) { best = n1; piv = 1; }
        if (n2 > best) { best = n2; piv = 2; }
        if (best < thr || best == 0.0) {
            rank = 0;
        } else {
            if (piv == 1) { swap_cols<0, 1>(A); int t = p0; p0 = p1; p1 = t; }
            if (piv == 2) { swap_cols<0, 2>(A); int t = p0; p0 = p2; p2 = t; }
            hh_step<0>(A, b);
        }
    }
    if (rank == 3) {  // k = 1
        n1 = colnorm2<1, 1>(A);
        n2 = colnorm2<1, 2>(A);
        if (fmax(n1, n2) < thr || fmax(n1, n2) == 0.0) {
            rank = 1;
        } else {
            if (n2 > n1) { swap_cols<1, 2>(A); int t = p1; p1 = p2; p2 = t; }
            hh_step<1>(A, b);
        }
    }
    if (rank == 3) {  // k = 2
        n2 = colnorm2<2, 2>(A);
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
                                              double plane_thres, double (&J)[6], double& ad, unsigned& n_cand) {
    const double px = sp.x, py = sp.y, pz = sp.z;
    const float qx = xform_row_d(pose[0], pose[1], pose[2], pose[9], px, py, pz);
    const float qy = xform_row_d(pose[3], pose[4], pose[5], pose[10], px, py, pz);
    const float qz = xform_row_d(pose[6], pose[7], pose[8], pose[11], px, py, pz);
    Knn5 nn;
    ivox_knn5_lists(map, qx, qy, qz, nn, n_cand);
    if (nn.j4 == 0xffffffffu) return false;  // fewer than 5 neighbours (:271-273)
    double A[5][3];
    {
        const float4 a0 = __ldg(map.lists + nn.j0), a1 = __ldg(map.lists + nn.j1), a2 = __ldg(map.lists + nn.j2),
                     a3 = __ldg(map.lists + nn.j3), a4 = __ldg(map.lists + nn.j4);
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

// columns of the per-point staging record
constexpr int kRecAd = 6, kRecValid = 7, kRecCand = 8, kRecHits = 9, kRecOne = 10, kRecW = 12;

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) p2plane_gn_kernel(P2PlaneLoopArgs a) {
    cg::grid_group grid = cg::this_grid();
    constexpr int W = BLOCK / 32;
    __shared__ double s_pose[12];
    __shared__ double s_rec[W][32][kRecW];
    __shared__ double s_red[W][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // which product of record columns lane k accumulates: sum_k = sgn * sum_p rec[p][ca] * rec[p][cb]
    int ca = kRecOne, cb = kRecOne;
    double sgn = 1.0;
    if (lane < 21) {
        int r = 0, k = lane;
        while (k >= 6 - r) { k -= 6 - r; ++r; }
        ca = r;
        cb = r + k;
    } else if (lane < 27) {
        ca = lane - 21; cb = kRecAd; sgn = -1.0;  // g = sum -J |d|
    } else if (lane == kAccValid) {
        ca = kRecValid; cb = kRecOne;
    } else if (lane == kAccRes) {
        ca = kRecAd; cb = kRecOne;
    } else if (lane == kAccCand) {
        ca = kRecCand; cb = kRecOne;
    } else if (lane == kAccHits) {
        ca = kRecHits; cb = kRecOne;
    } else {
        sgn = 0.0;
    }

    for (int it = 0; it < a.gp.max_iterations; ++it) {
        if (threadIdx.x < 9) s_pose[threadIdx.x] = __ldcg(&a.state->R[threadIdx.x]);
        else if (threadIdx.x < 12) s_pose[threadIdx.x] = __ldcg(&a.state->t[threadIdx.x - 9]);
        __syncthreads();

        double acc = 0.0;
        for (int base = blockIdx.x * BLOCK; base < a.n; base += gridDim.x * BLOCK) {
            const int i = base + threadIdx.x;
            double J[6] = {0, 0, 0, 0, 0, 0}, ad = 0.0, vflag = 0.0;
            unsigned n_cand = 0;
            if (i < a.n) {
                const float4 sp = a.src[i];
                bool use = p2plane_point(a.map, sp, s_pose, a.plane_thres, J, ad, n_cand);
                if (use) {
                    a.rec0[i] = make_float4((float)J[0], (float)J[1], (float)J[2], (float)J[3]);
                    a.rec1[i] = make_float4((float)J[4], (float)J[5], (float)ad, 1.0f);
                    a.flags[i] = 1;
                } else if (a.flags[i]) {  // stale contribution [quirk 1]
                    const float4 r0 = a.rec0[i], r1 = a.rec1[i];
                    J[0] = r0.x; J[1] = r0.y; J[2] = r0.z; J[3] = r0.w; J[4] = r1.x; J[5] = r1.y;
                    ad = r1.z;
                    use = true;
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) J[k] = 0.0;
                    ad = 0.0;
                }
                vflag = use ? 1.0 : 0.0;
            }
            double* rec = s_rec[warp][lane];
#pragma unroll
            for (int k = 0; k < 6; ++k) rec[k] = J[k];
            rec[kRecAd] = ad;
            rec[kRecValid] = vflag;
            rec[kRecCand] = (double)n_cand;
            rec[kRecHits] = (n_cand > 0) ? 1.0 : 0.0;
            rec[kRecOne] = 1.0;
            __syncwarp();
#pragma unroll 8
            for (int p = 0; p < 32; ++p) acc += s_rec[warp][p][ca] * s_rec[warp][p][cb];
            __syncwarp();
        }
        s_red[warp][lane] = acc * sgn;
        __syncthreads();
        if (threadIdx.x < kNumAcc) {
            double v = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) v += s_red[w][threadIdx.x];
            a.partials[(size_t)blockIdx.x * kAccStride + threadIdx.x] = v;
        }
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0) {
            // fixed-order cross-block reduction: W groups of 32 lanes stride over the rows, then groups 0..W-1
            double v = 0;
            if (lane < kNumAcc)
                for (int b = warp; b < (int)gridDim.x; b += W) v += __ldcg(&a.partials[(size_t)b * kAccStride + lane]);
            s_red[warp][lane] = v;
            __syncthreads();
            if (warp == 0) {
                double t = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) t += s_red[w][lane];
                __syncwarp();
                s_red[0][lane] = t;
                __syncwarp();
                if (lane == 0) {
                    gn_step(a.state, s_red[0], a.gp, a.log, a.log_cap);
                    __threadfence();
                }
            }
        }
        grid.sync();
        if (__ldcg(&a.state->done)) break;
    }
}

// ---- query ordering --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// 30-bit Morton code of the query's voxel at the initial pose (low 10 bits per axis: the 512-voxel period exceeds
// any scan's extent, and aliasing would only cost locality, never correctness)
__global__ void sortkey_kernel(const float4* __restrict__ src, int n, const GnState* __restrict__ state, float inv_res,
                               unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 sp = src[i];
    const double* R = state->R;
    const double* t = state->t;
    const float qx = xform_row_d(R[0], R[1], R[2], t[0], sp.x, sp.y, sp.z);
    const float qy = xform_row_d(R[3], R[4], R[5], t[1], sp.x, sp.y, sp.z);
    const float qz = xform_row_d(R[6], R[7], R[8], t[2], sp.x, sp.y, sp.z);
    keys[i] = spread10((unsigned)ivox_coord(qx, inv_res)) | (spread10((unsigned)ivox_coord(qy, inv_res)) << 1) |
              (spread10((unsigned)ivox_coord(qz, inv_res)) << 2);
    idx[i] = (unsigned)i;
}

__global__ void gather4_kernel(const float4* __restrict__ src, const unsigned* __restrict__ idx, int n, float4* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void ivox_knn_test_kernel(IvoxView map, const float4* __restrict__ q, int n, float4* __restrict__ out, int* __restrict__ found) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = q[i];
    Knn5 nn;
    unsigned nc;
    ivox_knn5_lists(map, p.x, p.y, p.z, nn, nc);
    const unsigned js[5] = {nn.j0, nn.j1, nn.j2, nn.j3, nn.j4};
    int f = 0;
    for (int k = 0; k < 5; ++k) {
        if (js[k] != 0xffffffffu) {
            out[(size_t)i * 5 + k] = map.lists[js[k]];
            ++f;
        } else {
            out[(size_t)i * 5 + k] = make_float4(0, 0, 0, 0);
        }
    }
    found[i] = f;
}

}  // namespace

int p2plane_max_grid(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p2plane_gn_kernel<kP2PlaneBlock>, kP2PlaneBlock, 0);
    const int g = sms * (per_sm > 0 ? per_sm : 1);
    if (device >= 0 && device < 64) cached[device] = g;
    return g;
}

int p2plane_grid(int n, int device) {
    const int need = (n + kP2PlaneBlock - 1) / kP2PlaneBlock;
    const int cap = p2plane_max_grid(device);
    const int g = need < cap ? need : cap;
    return g > 0 ? g : 1;
}

void launch_p2plane_loop(const P2PlaneLoopArgs& a, int grid, cudaStream_t st) {
    P2PlaneLoopArgs args = a;
    void* params[] = {&args};
    FLS_CUDA(cudaLaunchCooperativeKernel((const void*)p2plane_gn_kernel<kP2PlaneBlock>, dim3(grid), dim3(kP2PlaneBlock), params, 0, st));
}

// Morton-order the scan by the voxel each point falls into at the initial pose (state must be initialised).
void sort_queries(const float4* d_src, int n, const GnState* d_state, float inv_res, float4* d_sorted, BuildScratch& sc, cudaStream_t st,
                  int* launches) {
    if (n <= 0) return;
    sc.k32a.reserve(n);
    sc.k32b.reserve(n);
    sc.idx.reserve(n);
    sc.idx_sorted.reserve(n);
    sortkey_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_src, n, d_state, inv_res, sc.k32a.p, sc.idx.p);
    size_t t1 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, sc.k32a.p, sc.k32b.p, sc.idx.p, sc.idx_sorted.p, n, 0, 30, st);
    sc.cub_tmp.reserve(t1 + 256);
    size_t tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortPairs(sc.cub_tmp.p, tb, sc.k32a.p, sc.k32b.p, sc.idx.p, sc.idx_sorted.p, n, 0, 30, st));
    gather4_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_src, sc.idx_sorted.p, n, d_sorted);
    if (launches) *launches += 6;
}

void launch_ivox_knn_test(const IvoxView& map, const float4* d_q, int n, float4* d_out, int* d_found, cudaStream_t st) {
    if (n <= 0) return;
    ivox_knn_test_kernel<<<(n + 127) / 128, 128, 0, st>>>(map, d_q, n, d_out, d_found);
}

}  // namespace fls
