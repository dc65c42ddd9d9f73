// fls_gn.cu — K6 as stand-alone kernels: state initialisation and the per-iteration cross-block reduction +
// 6x6 solve + SE(3) update + stop rule, so a whole Match needs one device->host copy at the end instead of one
// per iteration.  (The LOAM path fuses the same step into its persistent kernel, fls_p2plane.cu.)
#include "fls_gn.cuh"

namespace fls {
namespace {

__global__ void gn_init_kernel(GnState* s, double t00, double t10, double t20, double t01, double t11, double t21, double t02, double t12,
                               double t22, double t03, double t13, double t23, int* sync, int n_sync) {
    for (int k = threadIdx.x; k < n_sync; k += blockDim.x) sync[k] = 0;
    if (threadIdx.x != 0) return;
    // arguments are the column-major Mat4d entries T(r,c) named t<r><c>
    const double R[9] = {t00, t01, t02, t10, t11, t12, t20, t21, t22};
    const double t[3] = {t03, t13, t23};
    for (int i = 0; i < 9; ++i) s->R[i] = s->R0[i] = s->Rprev[i] = R[i];
    for (int i = 0; i < 3; ++i) s->t[i] = s->t0[i] = s->tprev[i] = t[i];
    s->last_rot = s->last_pos = 0.0;
    for (int i = 0; i < 36; ++i) s->H[i] = 0;
    for (int i = 0; i < 6; ++i) s->g[i] = s->dx[i] = 0;
    s->sum_res = 0;
    s->cand_total = s->hits_total = 0;
    s->n_valid = 0;
    s->iter = 0;
    s->done = 0;
    s->converged = 0;
    s->failed = 0;
}

__global__ void __launch_bounds__(32) gn_solve_kernel(GnState* s, const double* __restrict__ partials, GnParams p, fls_iter_log* log,
                                                      int log_cap) {
    if (s->done) return;
    __shared__ double tot[kAccStride];
    const int lane = threadIdx.x;
    if (lane < kNumAcc) {
        double v = 0;
        for (int b = 0; b < p.n_blocks; ++b) v += partials[(size_t)b * kAccStride + lane];  // fixed order
        tot[lane] = v;
    }
    __syncwarp();
    if (lane == 0) gn_step(s, tot, p, log, log_cap);
}

}  // namespace

void launch_gn_init(GnState* d_state, const double* T, cudaStream_t st, int* d_sync, int n_sync) {
    // T is column-major: T[c*4 + r]
    gn_init_kernel<<<1, 128, 0, st>>>(d_state, T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10], T[12], T[13], T[14], d_sync, n_sync);
}

void launch_gn_solve(GnState* d_state, const double* d_partials, const GnParams& p, fls_iter_log* d_log, int log_capacity, cudaStream_t st) {
    gn_solve_kernel<<<1, 32, 0, st>>>(d_state, d_partials, p, d_log, log_capacity);
}

}  // namespace fls
