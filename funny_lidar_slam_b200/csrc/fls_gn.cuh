// fls_gn.cuh — Gauss-Newton plumbing shared by the residual kernels: the per-block reduction of the
// accumulators and the device-side solve / pose update / stop rule (K6).
//
// Conventions per plug-in (all line numbers upstream):
//   LOAM p2plane : dx=[dθ,dt], R <- Exp(dθ)·R, full-pivot solve, stop on thresholds OR |Δ‖dx‖|<1e-4,
//                  fail when n_valid < 50                      (loam_point_to_plane_ivox.h:167-203)
//   NDT          : dx=[dθ,dt], R <- R·Exp(dθ), H^-1·err, stop on thresholds, result forced true,
//                  early-out false when effective < min        (incremental_ndt.h:306-325)
//   ICP          : dx=[dt,dθ], R <- R·Exp(dθ), det==0 -> skip, converged only if thresholds met
//                                                              (icp_optimized.h:129-149)
#pragma once
#include "fls_common.cuh"

namespace fls {

struct GnParams {
    int method;  // fls_method: selects dx layout, update side, solver, stop rule
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

// One Gauss-Newton step from the reduced totals `tot[kNumAcc]`: fills H/g, solves, updates the pose in `s`,
// applies the plug-in's stop rule.  Executed by a single thread: every global read is issued up front (one L2
// round trip instead of ~30 serialized ones), the arithmetic runs on locals, the results are stored at the end.
__device__ inline void gn_step(GnState* s, const double* tot, const GnParams& p, fls_iter_log* log, int log_cap, int* release_flag = nullptr,
                               int release_value = 0) {
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldcg(&s->R[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = __ldcg(&s->t[i]);
    const double last_rot = __ldcg(&s->last_rot), last_pos = __ldcg(&s->last_pos);
    const double cand0 = __ldcg(&s->cand_total), hits0 = __ldcg(&s->hits_total);
    const int it = __ldcg(&s->iter);

    double H[36], g[6], dx[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) H[r * 6 + c] = H[c * 6 + r] = tot[tri6(r, c)];
    for (int a = 0; a < 6; ++a) g[a] = tot[21 + a];
    const long long n_valid = (long long)(tot[kAccValid] + 0.5);
    const double sum_res = tot[kAccRes];

    // pose before the update (LOAM-iVox map insertion rule)
#pragma unroll
    for (int i = 0; i < 9; ++i) s->Rprev[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) s->tprev[i] = t[i];

    bool stop = false;
    int converged = -1, failed = 0;
    double new_last_rot = last_rot, new_last_pos = last_pos;
    if (p.method == FLS_NDT && n_valid < (long long)p.min_effective) {
        failed = 1;  // incremental_ndt.h:306-309 — T = pose, return false
        converged = 0;
        stop = true;
    } else {
        double Rd[9], Rn[9];
        // H = sum J^T (W) J is symmetric positive definite whenever the problem is well posed: register-resident LDL^T
        // first, the pivoting solver that mirrors the reference's rank behaviour only when that is not safely the case.
        double det_spd = 0.0;
        const bool spd = solve6_spd(H, g, dx, &det_spd);
        if (p.method == FLS_ICP_P2P) {
            const double det = spd ? det_spd : solve6_lu(H, g, dx);
            if (det == 0.0) {
                for (int i = 0; i < 6; ++i) dx[i] = 0;  // icp_optimized.h:129-131 `continue`
            } else {
                for (int a = 0; a < 3; ++a) t[a] += dx[a];
                so3_exp(dx + 3, Rd);
                mat3_mul(R, Rd, Rn);
                for (int i = 0; i < 9; ++i) R[i] = Rn[i];
                if (norm3(dx + 3) < p.rot_thres && norm3(dx) < p.pos_thres) {
                    converged = 1;
                    stop = true;
                }
            }
        } else if (p.method == FLS_NDT) {
            if (!spd) solve6_lu(H, g, dx);
            so3_exp(dx, Rd);
            mat3_mul(R, Rd, Rn);
            for (int i = 0; i < 9; ++i) R[i] = Rn[i];
            for (int a = 0; a < 3; ++a) t[a] += dx[3 + a];
            if (norm3(dx) < p.rot_thres && norm3(dx + 3) < p.pos_thres) stop = true;
            converged = 1;  // forced true after the loop (incremental_ndt.h:325)
        } else {
            if (!spd) solve6_fullpiv(H, g, dx);
            so3_exp(dx, Rd);
            mat3_mul(Rd, R, Rn);
            for (int i = 0; i < 9; ++i) R[i] = Rn[i];
            for (int a = 0; a < 3; ++a) t[a] += dx[3 + a];
            const double rn = norm3(dx), pn = norm3(dx + 3);
            const double drot = fabs(rn - last_rot), dpos = fabs(pn - last_pos);
            new_last_rot = rn;
            new_last_pos = pn;
            if ((rn < p.rot_thres && pn < p.pos_thres) || (drot < 1.0e-4 && dpos < 1.0e-4)) stop = true;
            converged = (n_valid >= (long long)p.min_effective) ? 1 : 0;  // :201-203
        }
        if (it + 1 >= p.max_iterations) stop = true;
    }
    // ---- publish: what the other CTAs wait for goes out first (pose + done), then the hand-over flag, then the rest
#pragma unroll
    for (int i = 0; i < 9; ++i) s->R[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) s->t[i] = t[i];
    if (stop) s->done = 1;
    if (release_flag) {
        __threadfence();
        atomicExch(release_flag, release_value);
    }
    s->last_rot = new_last_rot;
    s->last_pos = new_last_pos;
    for (int i = 0; i < 36; ++i) s->H[i] = H[i];
    for (int i = 0; i < 6; ++i) {
        s->g[i] = g[i];
        s->dx[i] = dx[i];
    }
    s->n_valid = n_valid;
    s->sum_res = sum_res;
    s->cand_total = cand0 + tot[kAccCand];
    s->hits_total = hits0 + tot[kAccHits];
    s->iter = it + 1;
    if (converged >= 0) s->converged = converged;
    if (failed) s->failed = 1;
    if (log && it < log_cap) {
        fls_iter_log& L = log[it];
        for (int i = 0; i < 36; ++i) L.H[i] = H[i];
        for (int i = 0; i < 6; ++i) {
            L.g[i] = g[i];
            L.dx[i] = dx[i];
        }
        L.sum_residual = sum_res;
        L.n_valid = n_valid;
    }
}
#endif

}  // namespace fls
