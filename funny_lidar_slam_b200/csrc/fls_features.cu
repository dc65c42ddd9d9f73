// fls_features.cu — K4: LOAM edge / planar feature extraction on the projector's arrays.
//
// loam::FeatureExtractor::ExtractFeatures (src/loam/feature_extractor.cpp:35-222 upstream) =
//   SelectValidPoints (:64-118)  — occlusion / parallel-beam masks          -> feat_point_kernel (one thread per point)
//   ComputeRoughness  (:46-61)   — 11-tap range curvature                   -> feat_point_kernel
//   SelectFeatures    (:120-222) — per ring, 6 blocks: sort by roughness, greedy corner pick (<= 20, suppress +-5),
//                                  greedy planar suppression, emit every non-corner  -> feat_ring_kernel (one CTA per ring)
// Rings are independent (index gap 11 between rows, pointcloud_projector.cpp:115,131); blocks of a ring are not
// (suppression and the inclusive `block_end` visit bleed into the next block [quirk 10]), so a CTA walks its 6 blocks in
// order: bitonic sort of (roughness, position) in shared memory by the whole CTA — the composite key makes the unstable
// upstream std::sort deterministic exactly as the oracle pins it — then the two greedy passes by one thread on
// shared-memory state (they are sequential by definition: every pick changes the validity of later candidates).
#include <cstring>
#include <vector>

#include "fls_maps.h"

namespace fls {
namespace {

// per-point meta byte: bit0 valid, bit1 corner, bits 2-4 forward suppression reach, bits 5-7 backward reach
__device__ __forceinline__ int col_gap_ok(const int* __restrict__ col, int a, int b) { return abs(col[a] - col[b]) <= 10; }

__global__ void feat_point_kernel(const float* __restrict__ depth, const int* __restrict__ col, int n, float* __restrict__ rough,
                                  unsigned char* __restrict__ meta) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // roughness (:46-61): left-to-right fp32 sum, then - 10*d, squared; no FMA contraction
    float r = 0.f;
    if (k >= 5 && k < n - 5) {
        float s = __fadd_rn(depth[k - 5], depth[k - 4]);
        s = __fadd_rn(s, depth[k - 3]);
        s = __fadd_rn(s, depth[k - 2]);
        s = __fadd_rn(s, depth[k - 1]);
        s = __fadd_rn(s, depth[k + 1]);
        s = __fadd_rn(s, depth[k + 2]);
        s = __fadd_rn(s, depth[k + 3]);
        s = __fadd_rn(s, depth[k + 4]);
        s = __fadd_rn(s, depth[k + 5]);
        s = __fsub_rn(s, __fmul_rn(10.0f, depth[k]));
        r = __fmul_rn(s, s);
    }
    rough[k] = r;
    // validity (:64-118) in gather form: k is cleared by an occlusion found at i in [k, k+5] (near side, clears i-5..i),
    // by one found at i in [k-6, k-1] (far side, clears i+1..i+6), or by the parallel-beam test at k itself
    bool valid = !(k < 5 || k >= n - 6);
    const int lo = 5, hi = n - 6;  // i ranges over [lo, hi)
    if (valid) {
        for (int i = max(k, lo); i <= k + 5 && i < hi; ++i) {
            if (abs(col[i + 1] - col[i]) < 10 && (double)__fsub_rn(depth[i], depth[i + 1]) > 0.3) valid = false;
        }
        for (int i = max(k - 6, lo); i <= k - 1 && i < hi; ++i) {
            if (abs(col[i + 1] - col[i]) < 10 && !((double)__fsub_rn(depth[i], depth[i + 1]) > 0.3) &&
                (double)__fsub_rn(depth[i + 1], depth[i]) > 0.3)
                valid = false;
        }
        if (k >= lo && k < hi) {
            const float d1 = fabsf(__fsub_rn(depth[k - 1], depth[k])), d2 = fabsf(__fsub_rn(depth[k + 1], depth[k]));
            const double lim = 0.02 * (double)depth[k];
            if ((double)d1 > lim && (double)d2 > lim) valid = false;
        }
    } else if (k >= 5 && k < n - 6) {
        valid = false;
    }
    // points outside [5, n-6) are invalid but may still be cleared again harmlessly; occlusions can also clear them
    int fwd = 0, bwd = 0;
    if (k >= 5 && k < n - 5) {
        while (fwd < 5 && k + fwd + 1 < n && col_gap_ok(col, k + fwd + 1, k + fwd)) ++fwd;
        while (bwd < 5 && k - bwd - 1 >= 0 && col_gap_ok(col, k - bwd - 1, k - bwd)) ++bwd;
    }
    meta[k] = (unsigned char)((valid ? 1 : 0) | (fwd << 2) | (bwd << 5));
}

struct RingArgs {
    const float* __restrict__ rough;
    const unsigned char* __restrict__ meta;
    const int* __restrict__ row_start;
    const int* __restrict__ row_end;
    int n_rows;
    int n;
    float corner_thr, planar_thr;
    int lpad;  // power-of-two capacity of the sort buffer (>= longest block)
    int ring_cap;
    int* __restrict__ corner_out;  // [n_rows][120]
    int* __restrict__ corner_cnt;  // [n_rows]
    int* __restrict__ planar_out;  // ring r writes at planar_off[r]
    const int* __restrict__ planar_off;
    int* __restrict__ planar_cnt;
};

__device__ __forceinline__ void suppress(unsigned char* m, int li) {
    const int fwd = (m[li] >> 2) & 7, bwd = (m[li] >> 5) & 7;
    for (int k = 1; k <= fwd; ++k) m[li + k] &= ~1;
    for (int k = 1; k <= bwd; ++k) m[li - k] &= ~1;
}

__global__ void __launch_bounds__(256) feat_ring_kernel(RingArgs a) {
    extern __shared__ unsigned char smem[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem);  // [lpad] (roughness bits << 32 | point index)
    unsigned char* m = smem + (size_t)a.lpad * 8;                              // [ring_cap] meta bytes of this ring's points
    const int r = blockIdx.x;
    const int rs = a.row_start[r], re = a.row_end[r];
    const int p0 = rs - 5;             // first point of the ring
    const int ring_len = re + 6 - p0;  // points of the ring
    int n_corner = 0, n_planar = 0;
    int* pout = a.planar_out + a.planar_off[r];
    if (ring_len > 0 && ring_len <= a.ring_cap) {
        for (int i = threadIdx.x; i < ring_len; i += blockDim.x) m[i] = a.meta[p0 + i];
        __syncthreads();
        const int len = (re - rs) / 6;  // C integer division (:129)
        for (int b = 0; b < 6; ++b) {
            const int bs = rs + b * len, be = rs + (b + 1) * len;
            if (bs >= be) continue;
            const int L = be - bs;
            int lp = 1;
            while (lp < L) lp <<= 1;
            for (int i = threadIdx.x; i < lp; i += blockDim.x)
                skey[i] = (i < L) ? (((unsigned long long)__float_as_uint(a.rough[bs + i]) << 32) | (unsigned)(bs + i)) : ~0ull;
            __syncthreads();
            for (int k = 2; k <= lp; k <<= 1) {  // bitonic sort, ascending
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < lp; i += blockDim.x) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const unsigned long long x = skey[i], y = skey[ixj];
                            const bool up = ((i & k) == 0);
                            if ((x > y) == up) {
                                skey[i] = y;
                                skey[ixj] = x;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            if (threadIdx.x == 0) {
                // corner pass (:145-182): j = be (unsorted element of the next block) then the sorted block from the top
                int picked = 0;
                for (int j = be; j >= bs; --j) {
                    float rg;
                    int index;
                    if (j == be) {
                        rg = a.rough[be];
                        index = be;
                    } else {
                        const unsigned long long kv = skey[j - bs];
                        rg = __uint_as_float((unsigned)(kv >> 32));
                        index = (int)(unsigned)kv;
                        if (!(rg > a.corner_thr)) break;  // ascending order: nothing further down can qualify
                    }
                    const int li = index - p0;
                    if (rg > a.corner_thr && (m[li] & 1)) {
                        picked++;
                        if (picked <= 20) {
                            m[li] |= 2;
                            a.corner_out[r * 120 + n_corner++] = index;
                        } else {
                            break;
                        }
                        m[li] &= ~1;
                        suppress(m, li);
                    }
                }
                // planar pass (:184-217): ascending, inclusive of `be`; every non-corner is emitted
                for (int j = bs; j <= be; ++j) {
                    float rg;
                    int index;
                    if (j == be) {
                        rg = a.rough[be];
                        index = be;
                    } else {
                        const unsigned long long kv = skey[j - bs];
                        rg = __uint_as_float((unsigned)(kv >> 32));
                        index = (int)(unsigned)kv;
                    }
                    const int li = index - p0;
                    if ((m[li] & 1) && rg < a.planar_thr) {
                        m[li] &= ~1;
                        suppress(m, li);
                    }
                    if (!(m[li] & 2)) pout[n_planar++] = index;
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        a.corner_cnt[r] = n_corner;
        a.planar_cnt[r] = n_planar;
    }
}

__global__ void ring_offsets_kernel(const int* __restrict__ row_start, const int* __restrict__ row_end, int n_rows, int* __restrict__ off) {
    // capacity of ring r's planar segment: 6 * (len + 1)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int r = 0; r < n_rows; ++r) {
            off[r] = acc;
            const int len = (row_end[r] - row_start[r]) / 6;
            acc += len > 0 ? 6 * (len + 1) : 0;
        }
        off[n_rows] = acc;
    }
}

}  // namespace

// Host driver.  Returns fls_status; fills corner_idx / planar_idx (host) in the reference's emission order.
int extract_features_device(int device, const float* depth, const int* col, size_t n, const int* row_start, const int* row_end, int n_rows,
                            float corner_thr, float planar_thr, int* corner_idx, size_t* n_corner, int* planar_idx, size_t* n_planar,
                            fls_match_stats* stats) {
    *n_corner = 0;
    *n_planar = 0;
    if (n < 12 || n_rows <= 0) return FLS_OK;
    int max_len = 0, max_ring = 0;
    long long planar_cap = 0;
    for (int r = 0; r < n_rows; ++r) {
        const int len = (row_end[r] - row_start[r]) / 6;
        if (len > max_len) max_len = len;
        const int ring = row_end[r] + 6 - (row_start[r] - 5);
        if (ring > max_ring) max_ring = ring;
        if (len > 0 && (row_start[r] < 5 || row_end[r] + 6 > (int)n)) return FLS_ERR_INVALID_ARG;
        planar_cap += len > 0 ? 6LL * (len + 1) : 0;
    }
    int lpad = 1;
    while (lpad < max_len) lpad <<= 1;
    const size_t smem = (size_t)lpad * 8 + (size_t)(max_ring > 0 ? max_ring : 1) + 16;
    if (smem > 227 * 1024) return FLS_ERR_UNSUPPORTED;  // ring too long for the shared-memory working set (DESIGN.md "limits")
    FLS_CUDA(cudaSetDevice(device));
    cudaStream_t st;
    FLS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = FLS_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    try {
        FLS_CUDA(cudaEventCreate(&e0));
        FLS_CUDA(cudaEventCreate(&e1));
        DevBuf<float> d_depth, d_rough;
        DevBuf<int> d_col, d_rs, d_re, d_corner, d_ccnt, d_planar, d_poff, d_pcnt;
        DevBuf<unsigned char> d_meta;
        d_depth.reserve(n);
        d_rough.reserve(n);
        d_col.reserve(n);
        d_meta.reserve(n);
        d_rs.reserve(n_rows);
        d_re.reserve(n_rows);
        d_corner.reserve((size_t)n_rows * 120);
        d_ccnt.reserve(n_rows);
        d_planar.reserve((size_t)planar_cap + 8);
        d_poff.reserve(n_rows + 1);
        d_pcnt.reserve(n_rows);
        FLS_CUDA(cudaEventRecord(e0, st));
        FLS_CUDA(cudaMemcpyAsync(d_depth.p, depth, n * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(d_col.p, col, n * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(d_rs.p, row_start, n_rows * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(d_re.p, row_end, n_rows * 4, cudaMemcpyHostToDevice, st));
        feat_point_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_depth.p, d_col.p, (int)n, d_rough.p, d_meta.p);
        ring_offsets_kernel<<<1, 32, 0, st>>>(d_rs.p, d_re.p, n_rows, d_poff.p);
        RingArgs a;
        a.rough = d_rough.p;
        a.meta = d_meta.p;
        a.row_start = d_rs.p;
        a.row_end = d_re.p;
        a.n_rows = n_rows;
        a.n = (int)n;
        a.corner_thr = corner_thr;
        a.planar_thr = planar_thr;
        a.lpad = lpad;
        a.ring_cap = max_ring;
        a.corner_out = d_corner.p;
        a.corner_cnt = d_ccnt.p;
        a.planar_out = d_planar.p;
        a.planar_off = d_poff.p;
        a.planar_cnt = d_pcnt.p;
        FLS_CUDA(cudaFuncSetAttribute(feat_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        feat_ring_kernel<<<n_rows, 256, smem, st>>>(a);
        FLS_CUDA(cudaGetLastError());
        std::vector<int> h_corner((size_t)n_rows * 120), h_ccnt(n_rows), h_pcnt(n_rows), h_poff(n_rows + 1), h_planar((size_t)planar_cap + 8);
        FLS_CUDA(cudaMemcpyAsync(h_corner.data(), d_corner.p, h_corner.size() * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(h_ccnt.data(), d_ccnt.p, n_rows * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(h_pcnt.data(), d_pcnt.p, n_rows * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(h_poff.data(), d_poff.p, (n_rows + 1) * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(h_planar.data(), d_planar.p, (size_t)planar_cap * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaEventRecord(e1, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        size_t nc = 0, np = 0;
        for (int r = 0; r < n_rows; ++r) {
            for (int k = 0; k < h_ccnt[r]; ++k) corner_idx[nc++] = h_corner[(size_t)r * 120 + k];
            for (int k = 0; k < h_pcnt[r]; ++k) planar_idx[np++] = h_planar[(size_t)h_poff[r] + k];
        }
        *n_corner = nc;
        *n_planar = np;
        if (stats) {
            float ms = 0;
            FLS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            std::memset(stats, 0, sizeof(*stats));
            stats->gpu_ms = ms;
            stats->gpu_launches = 3;
            stats->n_source = (long long)n;
            stats->h2d_bytes = (long long)(n * 8 + (size_t)n_rows * 8);
            stats->d2h_bytes = (long long)((h_corner.size() + (size_t)planar_cap + 3 * (size_t)n_rows + 1) * 4);
        }
    } catch (const CudaError& e) {
        rc = e.status;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return rc;
}

}  // namespace fls
