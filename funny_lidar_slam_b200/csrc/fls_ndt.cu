// fls_ndt.cu — K2 (NDT residual + reduction) and K8-NDT (voxel map build / incremental update).
//
// IncrementalNDT (include/registration/incremental_ndt.h upstream) keeps an unordered_map of list nodes holding
// {points_, mu_, sigma_, information_, ndt_estimated_, num_points_}.  Here:
//   * table  : open-addressing slots {packed key (C-truncation of q/voxel, [quirk 5]), voxel index, estimated flag}
//   * hot    : 80-byte records {mu, symmetric information} — the only thing the Match kernel reads
//   * cold   : sigma, counters and a carry buffer of <= min_points_in_voxel pending points per voxel, touched only
//              by AddCloudToLocalMap / UpdateVoxel (:130-227)
// LRU eviction at `capacity` is not emulated (FLS_ERR_CAPACITY when the voxel count would reach it).
#include <cub/cub.cuh>

#include "fls_gn.cuh"
#include "fls_eig.cuh"
#include "fls_kernels.h"
#include "fls_maps.h"

namespace fls {
namespace {

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

__device__ __forceinline__ int ndt_coord(double v, double inv) { return (int)__dmul_rn(v, inv); }  // cast<int>() truncates

__global__ void ndt_keys_kernel(const float4* __restrict__ pts, size_t n, double inv, unsigned long long* __restrict__ keys,
                                unsigned* __restrict__ idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    keys[i] = pack_key(ndt_coord((double)p.x, inv), ndt_coord((double)p.y, inv), ndt_coord((double)p.z, inv));
    idx[i] = (unsigned)i;
}

__global__ void ndt_table_clear_kernel(HashSlot* tab, size_t slots, int* counter) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < slots) {
        tab[i].key = kEmptyKey;
        tab[i].start = 0xffffffffu;
        tab[i].count = 0;
    }
    if (i == 0) *counter = 0;
}

__device__ __forceinline__ void inv3_sym_reg(const double* S, double* Ai) {
    // cofactor inverse of a general 3x3 (Eigen's fixed-size inverse, incremental_ndt.h:134,151 upstream)
    const double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
    const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
    const double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = (S[2] * S[7] - S[1] * S[8]) * id; Ai[2] = (S[1] * S[5] - S[2] * S[4]) * id;
    Ai[3] = c01 * id; Ai[4] = (S[0] * S[8] - S[2] * S[6]) * id; Ai[5] = (S[2] * S[3] - S[0] * S[5]) * id;
    Ai[6] = c02 * id; Ai[7] = (S[1] * S[6] - S[0] * S[7]) * id; Ai[8] = (S[0] * S[4] - S[1] * S[3]) * id;
}

struct NdtUpdateArgs {
    const float4* __restrict__ pts;          // filtered cloud, map frame
    const unsigned* __restrict__ idx_sorted;  // point indices grouped by voxel (stable => cloud order inside a voxel)
    const unsigned long long* __restrict__ run_keys;
    const unsigned* __restrict__ starts;
    const unsigned* __restrict__ counts;
    int runs;
    HashSlot* tab;
    unsigned mask;
    NdtHot* hot;
    NdtCold* cold;
    double* carry;
    int* counter;
    long long capacity;
    int min_pts, max_pts;
    int first_scan;
};

// sequential mean / covariance over the carried points followed by the run's points (incremental_ndt.h:92-110)
__device__ void mean_cov_seq(const double* carry, int nc, const float4* __restrict__ pts, const unsigned* __restrict__ idx, unsigned s, unsigned c,
                             double* mean, double* cov) {
    const int n = nc + (int)c;
    double sm[3] = {0, 0, 0};
    for (int k = 0; k < nc; ++k)
        for (int a = 0; a < 3; ++a) sm[a] += carry[k * 3 + a];
    for (unsigned k = 0; k < c; ++k) {
        const float4 p = __ldg(pts + idx[s + k]);
        sm[0] += (double)p.x;
        sm[1] += (double)p.y;
        sm[2] += (double)p.z;
    }
    for (int a = 0; a < 3; ++a) mean[a] = sm[a] / (double)n;
    double cc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nc; ++k) {
        const double v[3] = {carry[k * 3] - mean[0], carry[k * 3 + 1] - mean[1], carry[k * 3 + 2] - mean[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) cc[a * 3 + b] += v[a] * v[b];
    }
    for (unsigned k = 0; k < c; ++k) {
        const float4 p = __ldg(pts + idx[s + k]);
        const double v[3] = {(double)p.x - mean[0], (double)p.y - mean[1], (double)p.z - mean[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) cc[a * 3 + b] += v[a] * v[b];
    }
    for (int a = 0; a < 9; ++a) cov[a] = cc[a] / (double)(n - 1);
}

__device__ __forceinline__ void store_hot(NdtHot* h, const double* mu, const double* info) {
    h->mu[0] = mu[0]; h->mu[1] = mu[1]; h->mu[2] = mu[2];
    h->info[0] = info[0];
    h->info[1] = 0.5 * (info[1] + info[3]);
    h->info[2] = 0.5 * (info[2] + info[6]);
    h->info[3] = info[4];
    h->info[4] = 0.5 * (info[5] + info[7]);
    h->info[5] = info[8];
    h->pad = 0.0;
}

// one thread per touched voxel: find-or-create, then UpdateVoxel (incremental_ndt.h:130-179 upstream)
__global__ void ndt_update_kernel(NdtUpdateArgs a, int* overflow) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.runs) return;
    const unsigned long long key = a.run_keys[r];
    const unsigned s = a.starts[r], c = a.counts[r];
    // find or insert
    unsigned h = hash_key(key) & a.mask;
    unsigned vi;
    bool created = false;
    for (;;) {
        const unsigned long long prev = atomicCAS(&a.tab[h].key, kEmptyKey, key);
        if (prev == kEmptyKey) {
            const int id = atomicAdd(a.counter, 1);
            if ((long long)id + 1 >= a.capacity) {  // data_.size() >= capacity_ would evict the LRU tail upstream (:203-206)
                atomicExch(overflow, 1);
            }
            vi = (unsigned)id;
            a.tab[h].start = vi;
            created = true;
            break;
        }
        if (prev == key) {
            vi = a.tab[h].start;
            break;
        }
        h = (h + 1) & a.mask;
    }
    if ((long long)vi >= a.capacity) return;
    NdtCold& cd = a.cold[vi];
    double* carry = a.carry + (size_t)vi * a.min_pts * 3;
    if (created) {
        cd.num_points = 0;
        cd.carry_count = 0;
        cd.estimated = 0;
        for (int k = 0; k < 9; ++k) cd.sigma[k] = 0;
    }
    if (!cd.estimated) cd.num_points += (int)c;  // VoxelData ctor / AddPoint (:66-76): counted only before the first estimate

    double mu[3], info[9];
    if (a.first_scan) {  // :131-143 — every touched voxel is (re-)estimated from the points of this call only
        if (c > 1u) {
            mean_cov_seq(carry, 0, a.pts, a.idx_sorted, s, c, mu, cd.sigma);
            double S[9];
            for (int k = 0; k < 9; ++k) S[k] = cd.sigma[k] + ((k % 4 == 0) ? 1.0e-3 : 0.0);
            inv3_sym_reg(S, info);
        } else {
            const float4 p = a.pts[a.idx_sorted[s]];
            mu[0] = (double)p.x; mu[1] = (double)p.y; mu[2] = (double)p.z;
            for (int k = 0; k < 9; ++k) info[k] = (k % 4 == 0) ? 1.0e2 : 0.0;
        }
        cd.estimated = 1;
        cd.carry_count = 0;
        store_hot(a.hot + vi, mu, info);
        a.tab[h].count = 1;
        return;
    }
    if (cd.estimated && cd.num_points > a.max_pts) return;  // :145-147 frozen
    const int nc = cd.carry_count;
    const int total = nc + (int)c;
    if (total > a.min_pts) {
        if (!cd.estimated) {  // :149-153
            mean_cov_seq(carry, nc, a.pts, a.idx_sorted, s, c, mu, cd.sigma);
            double S[9];
            for (int k = 0; k < 9; ++k) S[k] = cd.sigma[k] + ((k % 4 == 0) ? 1.0e-3 : 0.0);
            inv3_sym_reg(S, info);
            cd.estimated = 1;
        } else {  // :154-178 running merge + eigen clamp
            double cm[3], cv[9], nm[3], nv[9], om[3];
            mean_cov_seq(carry, nc, a.pts, a.idx_sorted, s, c, cm, cv);
            const NdtHot& oh = a.hot[vi];
            om[0] = oh.mu[0]; om[1] = oh.mu[1]; om[2] = oh.mu[2];
            const double m = (double)cd.num_points, n = (double)total;
            for (int k = 0; k < 3; ++k) nm[k] = (m * om[k] + n * cm[k]) / (m + n);
            for (int x = 0; x < 3; ++x)
                for (int y = 0; y < 3; ++y)
                    nv[x * 3 + y] = (m * (cd.sigma[x * 3 + y] + (om[x] - nm[x]) * (om[y] - nm[y])) + n * (cv[x * 3 + y] + (cm[x] - nm[x]) * (cm[y] - nm[y]))) / (m + n);
            for (int k = 0; k < 3; ++k) mu[k] = nm[k];
            for (int k = 0; k < 9; ++k) cd.sigma[k] = nv[k];
            cd.num_points += total;
            double lam[3], V[9];
            sym_eig3_dev(cd.sigma, lam, V);
            if (lam[1] < lam[0] * 1e-3) lam[1] = lam[0] * 1e-3;
            if (lam[2] < lam[0] * 1e-3) lam[2] = lam[0] * 1e-3;
            for (int x = 0; x < 3; ++x)
                for (int y = 0; y < 3; ++y)
                    info[x * 3 + y] = V[x * 3 + 0] * V[y * 3 + 0] / lam[0] + V[x * 3 + 1] * V[y * 3 + 1] / lam[1] + V[x * 3 + 2] * V[y * 3 + 2] / lam[2];
        }
        cd.carry_count = 0;
        store_hot(a.hot + vi, mu, info);
        a.tab[h].count = 1;
    } else {  // keep the points for a later estimate
        for (unsigned k = 0; k < c; ++k) {
            const float4 p = a.pts[a.idx_sorted[s + k]];
            carry[(nc + k) * 3 + 0] = (double)p.x;
            carry[(nc + k) * 3 + 1] = (double)p.y;
            carry[(nc + k) * 3 + 2] = (double)p.z;
        }
        cd.carry_count = total;
    }
}

// ---- K2: NDT residual kernel -----------------------------------------------------------------------------------
// One persistent launch runs every Gauss-Newton iteration of a Match (gn_handover, fls_gn.cuh): grid-stride over the
// points, per-thread sums, CTA row, last-CTA fold + solve + release.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) ndt_gn_kernel(NdtArgs a, GnLoopCtl ctl) {
    __shared__ double s_pose[12];
    if (threadIdx.x < 9) s_pose[threadIdx.x] = __ldcg(&a.state->R[threadIdx.x]);
    else if (threadIdx.x < 12) s_pose[threadIdx.x] = __ldcg(&a.state->t[threadIdx.x - 9]);
    __syncthreads();
    for (int it = 0; it < ctl.gp.max_iterations; ++it) {  // the hand-over leaves the next pose in s_pose
    double acc[kNumAcc];
#pragma unroll
    for (int k = 0; k < kNumAcc; ++k) acc[k] = 0.0;

    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < a.n; i += gridDim.x * BLOCK) {
        const float4 sp = a.src[i];
        const double px = sp.x, py = sp.y, pz = sp.z;
        const double* R = s_pose;
        const double qx = xform_row_dd(R[0], R[1], R[2], s_pose[9], px, py, pz);  // incremental_ndt.h:255
        const double qy = xform_row_dd(R[3], R[4], R[5], s_pose[10], px, py, pz);
        const double qz = xform_row_dd(R[6], R[7], R[8], s_pose[11], px, py, pz);
        const int kx = ndt_coord(qx, a.map.inv_voxel), ky = ndt_coord(qy, a.map.inv_voxel), kz = ndt_coord(qz, a.map.inv_voxel);  // :256
        double L0 = 0, L1 = 0, L2 = 0, L3 = 0, L4 = 0, L5 = 0;  // sum of information matrices (sym)
        double w0 = 0, w1 = 0, w2 = 0, chis = 0;
        int cnt = 0, hits = 0;
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {  // stencil order of :122-127 = first 7 entries of c_stencil
            const unsigned long long key = pack_key(kx + c_stencil[s][0], ky + c_stencil[s][1], kz + c_stencil[s][2]);
            unsigned vi, est;
            if (!table_find(a.map.tab, a.map.mask, key, vi, est)) continue;
            ++hits;
            if (!est) continue;  // voxel exists but ndt_estimated_ is false (:263)
            const double2* hp = reinterpret_cast<const double2*>(a.map.hot + vi);
            const double2 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2), h3 = __ldg(hp + 3), h4 = __ldg(hp + 4);
            const double ex = qx - h0.x, ey = qy - h0.y, ez = qz - h1.x;
            const double ixx = h1.y, ixy = h2.x, ixz = h2.y, iyy = h3.x, iyz = h3.y, izz = h4.x;
            const double ux = ixx * ex + ixy * ey + ixz * ez;
            const double uy = ixy * ex + iyy * ey + iyz * ez;
            const double uz = ixz * ex + iyz * ey + izz * ez;
            const double chi = ex * ux + ey * uy + ez * uz;
            if (isnan(chi) || chi > a.outlier_thres) continue;  // :267-271
            L0 += ixx; L1 += ixy; L2 += ixz; L3 += iyy; L4 += iyz; L5 += izz;
            w0 += ux; w1 += uy; w2 += uz;
            chis += chi;
            ++cnt;
        }
        acc[kAccHits] += (double)hits;
        acc[kAccCand] += (double)cnt;
        if (cnt > 0) {
            // B = -R * hat(p)  (3x3), J = [B | I]  (:273-275);  H = [[B^T L B, B^T L],[L B, L]],  err = -[B^T w ; w]
            double B[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double r0 = R[r * 3 + 0], r1 = R[r * 3 + 1], r2 = R[r * 3 + 2];
                B[r][0] = -(r1 * pz - r2 * py);
                B[r][1] = -(r2 * px - r0 * pz);
                B[r][2] = -(r0 * py - r1 * px);
            }
            double LB[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                LB[0][c] = L0 * B[0][c] + L1 * B[1][c] + L2 * B[2][c];
                LB[1][c] = L1 * B[0][c] + L3 * B[1][c] + L4 * B[2][c];
                LB[2][c] = L2 * B[0][c] + L4 * B[1][c] + L5 * B[2][c];
            }
            // upper triangle in dx = [dθ(0..2), dt(3..5)] order
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = r; c < 3; ++c) acc[tri6(r, c)] += B[0][r] * LB[0][c] + B[1][r] * LB[1][c] + B[2][r] * LB[2][c];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[tri6(r, 3 + c)] += LB[c][r];  // (B^T L)[r][c] = (L B)[c][r], L symmetric
            acc[tri6(3, 3)] += L0; acc[tri6(3, 4)] += L1; acc[tri6(3, 5)] += L2;
            acc[tri6(4, 4)] += L3; acc[tri6(4, 5)] += L4; acc[tri6(5, 5)] += L5;
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[21 + r] -= (B[0][r] * w0 + B[1][r] * w1 + B[2][r] * w2);
            acc[24] -= w0; acc[25] -= w1; acc[26] -= w2;
            acc[kAccValid] += (double)cnt;
            acc[kAccRes] += chis;
        }
    }
    if (gn_handover<BLOCK>(acc, ctl, it, s_pose)) break;
    }
}

}  // namespace

int ndt_grid(int n, int device) {
    static int cap[64] = {0};
    if (device >= 0 && device < 64 && !cap[device]) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ndt_gn_kernel<kNdtBlock>, kNdtBlock, 0);
        cap[device] = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int need = (n + kNdtBlock - 1) / kNdtBlock;
    const int c = (device >= 0 && device < 64) ? cap[device] : 148;
    const int g = need < c ? need : c;
    return g > 0 ? g : 1;
}
void launch_ndt_loop(const NdtArgs& a, const GnLoopCtl& ctl, int grid, cudaStream_t st) {
    NdtArgs a_ = a;
    GnLoopCtl c_ = ctl;
    void* params[] = {&a_, &c_};
    FLS_CUDA(cudaLaunchCooperativeKernel((const void*)ndt_gn_kernel<kNdtBlock>, dim3(grid), dim3(kNdtBlock), params, 0, st));
}

void NdtMap::configure(double voxel_size, int min_points, int max_points, long long cap) {
    voxel = voxel_size;
    inv_voxel = 1.0 / voxel_size;
    min_pts = min_points;
    max_pts = max_points;
    capacity = cap;
}

int NdtMap::add_cloud(const float4* d_cloud, size_t n, float leaf, bool first_scan, cudaStream_t st) {
    if (n == 0) return FLS_OK;
    filtered.reserve(n);
    int l = 0;
    const size_t nf = voxel_grid_device(d_cloud, n, leaf, filtered.p, scratch, st, &l);  // :186
    launches += l;
    if (nf == 0) return FLS_OK;
    if (slots == 0) {  // first use: size everything by the configured capacity
        size_t want = 1024;
        while (want < 2 * (size_t)capacity) want <<= 1;
        slots = want;
        mask = (unsigned)(slots - 1);
        table.reserve(slots);
        hot.reserve((size_t)capacity);
        cold.reserve((size_t)capacity);
        carry.reserve((size_t)capacity * (size_t)(min_pts > 0 ? min_pts : 1) * 3);
        counter.reserve(2);
        ndt_table_clear_kernel<<<grid_for(slots, 256), 256, 0, st>>>(table.p, slots, counter.p);
        launches++;
    }
    BuildScratch& sc = scratch;
    sc.keys.reserve(nf);
    sc.keys_sorted.reserve(nf);
    sc.uniq.reserve(nf);
    sc.idx.reserve(nf);
    sc.idx_sorted.reserve(nf);
    sc.counts.reserve(nf);
    sc.starts.reserve(nf);
    sc.num_runs.reserve(2);
    ndt_keys_kernel<<<grid_for(nf, 256), 256, 0, st>>>(filtered.p, nf, inv_voxel, sc.keys.p, sc.idx.p);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, sc.keys.p, sc.keys_sorted.p, sc.idx.p, sc.idx_sorted.p, (int)nf, 0, 63, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, t2, sc.keys_sorted.p, sc.uniq.p, sc.counts.p, sc.num_runs.p, (int)nf, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, sc.counts.p, sc.starts.p, (int)nf, st);
    size_t tmp = t1 > t2 ? t1 : t2;
    tmp = tmp > t3 ? tmp : t3;
    sc.cub_tmp.reserve(tmp + 256);
    size_t tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortPairs(sc.cub_tmp.p, tb, sc.keys.p, sc.keys_sorted.p, sc.idx.p, sc.idx_sorted.p, (int)nf, 0, 63, st));
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRunLengthEncode::Encode(sc.cub_tmp.p, tb, sc.keys_sorted.p, sc.uniq.p, sc.counts.p, sc.num_runs.p, (int)nf, st));
    FLS_CUDA(cudaMemcpyAsync(sc.h_num_runs, sc.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const int runs = *sc.h_num_runs;
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceScan::ExclusiveSum(sc.cub_tmp.p, tb, sc.counts.p, sc.starts.p, runs, st));
    NdtUpdateArgs ua;
    ua.pts = filtered.p;
    ua.idx_sorted = sc.idx_sorted.p;
    ua.run_keys = sc.uniq.p;
    ua.starts = sc.starts.p;
    ua.counts = sc.counts.p;
    ua.runs = runs;
    ua.tab = table.p;
    ua.mask = mask;
    ua.hot = hot.p;
    ua.cold = cold.p;
    ua.carry = carry.p;
    ua.counter = counter.p;
    ua.capacity = capacity;
    ua.min_pts = min_pts;
    ua.max_pts = max_pts;
    ua.first_scan = first_scan ? 1 : 0;
    int* overflow = counter.p + 1;
    FLS_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), st));
    ndt_update_kernel<<<grid_for(runs, 128), 128, 0, st>>>(ua, overflow);
    int h[2] = {0, 0};
    FLS_CUDA(cudaMemcpyAsync(h, counter.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    launches += 6;
    n_vox = (size_t)h[0];
    if (h[1]) return FLS_ERR_CAPACITY;
    return FLS_OK;
}

}  // namespace fls
