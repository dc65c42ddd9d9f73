// fls_gn.cuh — Gauss-Newton plumbing shared by the three residual kernels: the per-block reduction of the
// 29 accumulators and the launch interface of the device-side solve / pose update (K6).
#pragma once
#include "fls_common.cuh"

namespace fls {

struct GnParams {
    int method;  // fls_method: selects dx layout, update side, solver, stop rule (SURVEY.md §8a convention table)
    int max_iterations;
    int min_effective;  // NDT: min_effective_pts; LOAM: 50 valid planar points
    int n_blocks;       // rows of the partial-sum matrix produced by the residual kernel
    double rot_thres, pos_thres;
};

void launch_gn_init(GnState* d_state, const double* T_colmajor, cudaStream_t st);
void launch_gn_solve(GnState* d_state, const double* d_partials, const GnParams& p, fls_iter_log* d_log, int log_capacity, cudaStream_t st);

#ifdef __CUDACC__
// Reduce acc[kNumAcc] over the thread block (blockDim.x multiple of 32, <= 1024) and write one row of the
// partial-sum matrix.  Fixed order: lanes by xor-butterfly, then warps 0..W-1 => bitwise reproducible.
template <int BLOCK>
__device__ __forceinline__ void block_reduce_store(double (&acc)[kNumAcc], double* __restrict__ partial_row) {
    constexpr int W = BLOCK / 32;
    __shared__ double s_red[W][kAccStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNumAcc; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNumAcc) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) v += s_red[w][threadIdx.x];
        partial_row[threadIdx.x] = v;
    }
}
#endif

}  // namespace fls
