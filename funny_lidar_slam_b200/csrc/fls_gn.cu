// fls_gn.cu — K6: cross-block reduction, 6x6 solve, SE(3) update and stop rule on the device, so a whole
// Match needs one device->host copy at the end instead of one per iteration.
//
// Conventions per plug-in (all line numbers upstream):
//   LOAM p2plane : dx=[dθ,dt], R <- Exp(dθ)·R, full-pivot solve, stop on thresholds OR |Δ‖dx‖|<1e-4,
//                  fail when n_valid < 50                      (loam_point_to_plane_ivox.h:167-203)
//   NDT          : dx=[dθ,dt], R <- R·Exp(dθ), H^-1·err, stop on thresholds, result forced true,
//                  early-out false when effective < min        (incremental_ndt.h:306-325)
//   ICP          : dx=[dt,dθ], R <- R·Exp(dθ), det==0 -> skip, converged only if thresholds met
//                                                              (icp_optimized.h:129-149)
#include "fls_gn.cuh"

namespace fls {
namespace {

__global__ void gn_init_kernel(GnState* s, double t00, double t10, double t20, double t01, double t11, double t21, double t02, double t12,
                               double t22, double t03, double t13, double t23) {
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
    if (lane != 0) return;

    double H[36], g[6], dx[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) H[r * 6 + c] = H[c * 6 + r] = tot[tri6(r, c)];
    for (int a = 0; a < 6; ++a) g[a] = tot[21 + a];
    const long long n_valid = (long long)(tot[kAccValid] + 0.5);
    const double sum_res = tot[kAccRes];
    for (int i = 0; i < 36; ++i) s->H[i] = H[i];
    for (int i = 0; i < 6; ++i) s->g[i] = g[i];
    s->n_valid = n_valid;
    s->sum_res = sum_res;
    s->cand_total += tot[kAccCand];
    s->hits_total += tot[kAccHits];
    const int it = s->iter;
    s->iter = it + 1;
    for (int i = 0; i < 9; ++i) s->Rprev[i] = s->R[i];
    for (int i = 0; i < 3; ++i) s->tprev[i] = s->t[i];

    bool stop = false;
    if (p.method == FLS_NDT && n_valid < (long long)p.min_effective) {
        // incremental_ndt.h:306-309 — T = pose, return false
        s->failed = 1;
        s->converged = 0;
        s->done = 1;
        stop = true;
    } else {
        bool updated = true;
        double Rd[9], Rn[9];
        if (p.method == FLS_ICP_P2P) {
            const double det = solve6_lu(H, g, dx);
            if (det == 0.0) {
                updated = false;
                for (int i = 0; i < 6; ++i) dx[i] = 0;
            } else {
                for (int a = 0; a < 3; ++a) s->t[a] += dx[a];
                so3_exp(dx + 3, Rd);
                mat3_mul(s->R, Rd, Rn);
                for (int i = 0; i < 9; ++i) s->R[i] = Rn[i];
                if (norm3(dx + 3) < p.rot_thres && norm3(dx) < p.pos_thres) {
                    s->converged = 1;
                    stop = true;
                }
            }
        } else if (p.method == FLS_NDT) {
            solve6_lu(H, g, dx);
            so3_exp(dx, Rd);
            mat3_mul(s->R, Rd, Rn);
            for (int i = 0; i < 9; ++i) s->R[i] = Rn[i];
            for (int a = 0; a < 3; ++a) s->t[a] += dx[3 + a];
            if (norm3(dx) < p.rot_thres && norm3(dx + 3) < p.pos_thres) stop = true;
            s->converged = 1;  // forced true after the loop (incremental_ndt.h:325)
        } else {
            solve6_fullpiv(H, g, dx);
            so3_exp(dx, Rd);
            mat3_mul(Rd, s->R, Rn);
            for (int i = 0; i < 9; ++i) s->R[i] = Rn[i];
            for (int a = 0; a < 3; ++a) s->t[a] += dx[3 + a];
            const double rn = norm3(dx), pn = norm3(dx + 3);
            const double drot = fabs(rn - s->last_rot), dpos = fabs(pn - s->last_pos);
            s->last_rot = rn;
            s->last_pos = pn;
            if ((rn < p.rot_thres && pn < p.pos_thres) || (drot < 1.0e-4 && dpos < 1.0e-4)) stop = true;
            s->converged = (n_valid >= (long long)p.min_effective) ? 1 : 0;  // :201-203
        }
        (void)updated;
        for (int i = 0; i < 6; ++i) s->dx[i] = dx[i];
        if (it + 1 >= p.max_iterations) stop = true;
        if (stop) s->done = 1;
    }
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

}  // namespace

void launch_gn_init(GnState* d_state, const double* T, cudaStream_t st) {
    // T is column-major: T[c*4 + r]
    gn_init_kernel<<<1, 32, 0, st>>>(d_state, T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10], T[12], T[13], T[14]);
}

void launch_gn_solve(GnState* d_state, const double* d_partials, const GnParams& p, fls_iter_log* d_log, int log_capacity, cudaStream_t st) {
    gn_solve_kernel<<<1, 32, 0, st>>>(d_state, d_partials, p, d_log, log_capacity);
}

}  // namespace fls
