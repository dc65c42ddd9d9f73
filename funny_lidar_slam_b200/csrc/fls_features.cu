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
#include <cstdlib>
#include <cstring>
#include <mutex>
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
    int lpad;      // power-of-two capacity of the sort buffer (>= longest block)
    int lcap;      // longest block + 16: capacity of the block-local visit tables
    int ring_cap;  // longest ring
    int max_rounds;  // parallel rounds per greedy pass before one thread finishes the rest in visit order
    unsigned long long* __restrict__ sorted;  // [n] per-block ascending (roughness bits << 32 | point index), written by feat_sort_kernel
    int* __restrict__ corner_out;             // [n_rows][120]
    int* __restrict__ corner_cnt;             // [n_rows]
    int* __restrict__ planar_out;             // ring r writes at planar_off[r]
    const int* __restrict__ planar_off;
    int* __restrict__ planar_cnt;
};

// K4a: the six per-ring sorts do not depend on the greedy state, so every (ring, block) sorts in its own CTA.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) feat_sort_kernel(RingArgs a) {
    extern __shared__ unsigned char smem[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem);
    const int r = blockIdx.y, b = blockIdx.x;
    const int rs = a.row_start[r], re = a.row_end[r];
    const int len = (re - rs) / 6;  // C integer division (:129)
    const int bs = rs + b * len, be = rs + (b + 1) * len;
    if (bs >= be) return;
    const int L = be - bs;
    int lp = 1;
    while (lp < L) lp <<= 1;
    for (int i = threadIdx.x; i < lp; i += BLOCK)
        skey[i] = (i < L) ? (((unsigned long long)__float_as_uint(a.rough[bs + i]) << 32) | (unsigned)(bs + i)) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= lp; k <<= 1) {  // bitonic sort, ascending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < lp; i += BLOCK) {
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
    for (int i = threadIdx.x; i < L; i += BLOCK) a.sorted[bs + i] = skey[i];
}

enum : unsigned char { kDead = 0, kUndecided = 1, kPicked = 2 };

// exclusive prefix sum of one int per thread over the CTA; `total` = sum over all threads
template <int BLOCK>
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp /*[BLOCK/32]*/, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    return base + inc - v;
}

// The greedy passes of SelectFeatures visit candidates in a fixed order and a pick only suppresses points at most 5
// positions away, so the outcome of visit v depends only on earlier visits within +-5 positions whose reach covers it.
// Every CTA round decides all visits whose earlier neighbours are decided (the parallel form of a sequential greedy
// independent set: same picks, bit for bit); a monotone roughness ramp degenerates to one decision per round, so after
// `max_rounds` one thread finishes the rest in visit order.
//   vpos[v]  ring-local position of visit v,  rank[q] visit number of block-local position q (INT_MAX: not visited),
//   stat[q]  kDead / kUndecided / kPicked,     m[]     ring-local meta bytes (reach bits are constant during a pass)
template <int BLOCK>
__device__ __forceinline__ void greedy_rounds(int nvis, const int* vpos, const int* rank, volatile unsigned char* stat, const unsigned char* m,
                                              int base_li /*ring-local position of block-local 0*/, int span, int max_rounds) {
    auto decide = [&](int v) -> int {  // kDead / kPicked / kUndecided (= wait)
        const int li = vpos[v], q0 = li - base_li;
        bool wait = false, dead = false;
#pragma unroll
        for (int d = -5; d <= 5; ++d) {
            if (d == 0) continue;
            const int q = q0 + d;
            if (q < 0 || q >= span) continue;
            if (rank[q] >= v) continue;  // not visited before v
            const unsigned char mm = m[li + d];
            const int reach = d < 0 ? ((mm >> 2) & 7) : ((mm >> 5) & 7);  // neighbour before v: its forward reach; after: backward
            if ((d < 0 ? -d : d) > reach) continue;
            const unsigned char s = stat[q];
            dead |= (s == kPicked);
            wait |= (s == kUndecided);
        }
        return dead ? kDead : (wait ? kUndecided : kPicked);
    };
    int rounds = 0;
    bool more;
    do {
        bool und = false;
        for (int v = threadIdx.x; v < nvis; v += BLOCK) {
            const int q0 = vpos[v] - base_li;
            if (stat[q0] != kUndecided) continue;
            const int s = decide(v);
            if (s == kUndecided) und = true;
            else stat[q0] = (unsigned char)s;
        }
        more = __syncthreads_or(und);
    } while (more && ++rounds < max_rounds);
    if (more) {
        if (threadIdx.x == 0)
            for (int v = 0; v < nvis; ++v) {
                const int q0 = vpos[v] - base_li;
                if (stat[q0] == kUndecided) stat[q0] = (unsigned char)decide(v);  // every earlier visit is decided by now
            }
        __syncthreads();
    }
}

__device__ __forceinline__ void suppress(unsigned char* m, int li) {
    const int fwd = (m[li] >> 2) & 7, bwd = (m[li] >> 5) & 7;
    for (int k = 1; k <= fwd; ++k) m[li + k] &= ~1;
    for (int k = 1; k <= bwd; ++k) m[li - k] &= ~1;
}

// K4b: one CTA per ring walks its six blocks in order (suppression and the inclusive `block_end` visit bleed into the
// next block [quirk 10]); inside a block everything is CTA-parallel.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) feat_ring_kernel(RingArgs a) {
    extern __shared__ unsigned char smem[];
    int* rank = reinterpret_cast<int*>(smem);                                // [lcap]
    int* vpos = rank + a.lcap;                                               // [lcap]
    unsigned char* stat = reinterpret_cast<unsigned char*>(vpos + a.lcap);  // [lcap]
    unsigned char* m = stat + a.lcap;                                        // [ring_cap]
    __shared__ int s_warp[BLOCK / 32];
    __shared__ int s_cut;
    const int r = blockIdx.x;
    const int rs = a.row_start[r], re = a.row_end[r];
    const int p0 = rs - 5;             // first point of the ring
    const int ring_len = re + 6 - p0;  // points of the ring
    int n_corner = 0, n_planar = 0;
    int* pout = a.planar_out + a.planar_off[r];
    if (ring_len > 0 && ring_len <= a.ring_cap) {
        for (int i = threadIdx.x; i < ring_len; i += BLOCK) m[i] = a.meta[p0 + i];
        const int len = (re - rs) / 6;  // C integer division (:129)
        for (int b = 0; b < 6; ++b) {
            const int bs = rs + b * len, be = rs + (b + 1) * len;
            if (bs >= be) continue;
            const int L = be - bs;
            const int base = bs - 5, span = L + 11;  // block-local window: positions bs-5 .. be+5
            const int base_li = base - p0;
            const unsigned long long* sk = a.sorted + bs;
            const float rough_be = a.rough[be];

            // ---- corner pass (:145-182): visit `be` (unsorted element of the next block), then the sorted block from the top
            // while roughness > threshold; the 21st pick ends the pass untouched.
            if (threadIdx.x == 0) s_cut = -1;
            for (int q = threadIdx.x; q < span; q += BLOCK) rank[q] = 0x7fffffff;
            __syncthreads();
            {
                int cut = -1;  // highest sorted slot that fails `roughness > threshold` — the descending walk stops there
                for (int i = threadIdx.x; i < L; i += BLOCK)
                    if (!(__uint_as_float((unsigned)(sk[i] >> 32)) > a.corner_thr)) cut = i;
                if (cut >= 0) atomicMax(&s_cut, cut);
            }
            __syncthreads();
            const int first = s_cut + 1;        // sorted slots first..L-1 are visited, top first
            const int nvis_c = 1 + (L - first);  // visit 0 is `be`
            for (int v = threadIdx.x; v < nvis_c; v += BLOCK) {
                const int idx = (v == 0) ? be : (int)(unsigned)sk[L - v];
                const float rg = (v == 0) ? rough_be : __uint_as_float((unsigned)(sk[L - v] >> 32));
                const int li = idx - p0, q = li - base_li;
                vpos[v] = li;
                rank[q] = v;
                stat[q] = (rg > a.corner_thr && (m[li] & 1)) ? kUndecided : kDead;
            }
            __syncthreads();
            greedy_rounds<BLOCK>(nvis_c, vpos, rank, stat, m, base_li, span, a.max_rounds);
            {
                const int chunk = (nvis_c + BLOCK - 1) / BLOCK;
                const int v0 = threadIdx.x * chunk, v1 = min(nvis_c, v0 + chunk);
                int cnt = 0;
                for (int v = v0; v < v1; ++v) cnt += (stat[vpos[v] - base_li] == kPicked);
                int total;
                int pr = block_excl_scan<BLOCK>(cnt, s_warp, total);
                for (int v = v0; v < v1; ++v) {
                    const int li = vpos[v], q = li - base_li;
                    if (stat[q] != kPicked) continue;
                    if (pr < 20) {
                        m[li] = (unsigned char)((m[li] | 2) & ~1);  // own byte only; neighbours are cleared after the barrier
                        a.corner_out[r * 120 + n_corner + pr] = li + p0;
                    } else {
                        stat[q] = kDead;  // beyond the limit: never picked, state untouched
                    }
                    ++pr;
                }
                n_corner += min(total, 20);
                __syncthreads();
                for (int v = v0; v < v1; ++v) {
                    const int li = vpos[v];
                    if (stat[li - base_li] == kPicked) suppress(m, li);
                }
                __syncthreads();
            }

            // ---- planar pass (:184-217): ascending through the sorted block, then `be`; every visited non-corner is emitted
            const int nvis_p = L + 1;
            for (int v = threadIdx.x; v < nvis_p; v += BLOCK) {
                const int idx = (v == L) ? be : (int)(unsigned)sk[v];
                const float rg = (v == L) ? rough_be : __uint_as_float((unsigned)(sk[v] >> 32));
                const int li = idx - p0, q = li - base_li;
                vpos[v] = li;
                rank[q] = v;
                stat[q] = ((m[li] & 1) && rg < a.planar_thr) ? kUndecided : kDead;
            }
            __syncthreads();
            greedy_rounds<BLOCK>(nvis_p, vpos, rank, stat, m, base_li, span, a.max_rounds);
            {
                const int chunk = (nvis_p + BLOCK - 1) / BLOCK;
                const int v0 = threadIdx.x * chunk, v1 = min(nvis_p, v0 + chunk);
                for (int v = v0; v < v1; ++v) {
                    const int li = vpos[v];
                    if (stat[li - base_li] == kPicked) {
                        m[li] &= ~1;
                        suppress(m, li);  // every writer clears bit 0 only: order-free
                    }
                }
                int cnt = 0;
                for (int v = v0; v < v1; ++v) cnt += !(m[vpos[v]] & 2);  // corner bits are final since the barrier above
                int total;
                int off = block_excl_scan<BLOCK>(cnt, s_warp, total);
                for (int v = v0; v < v1; ++v)
                    if (!(m[vpos[v]] & 2)) pout[n_planar + off++] = vpos[v] + p0;
                n_planar += total;
                __syncthreads();
            }
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

// Per-device workspace: stream, events, device buffers and pinned staging survive across calls (the extractor runs
// once per scan; allocating per call cost more than the kernels).  Calls on one device are serialised by the mutex.
struct FeatWorkspace {
    std::mutex mu;
    bool ready = false;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, k0 = nullptr, k1 = nullptr;
    DevBuf<float> d_depth, d_rough;
    DevBuf<int> d_col, d_rows, d_out, d_poff;
    DevBuf<unsigned char> d_meta;
    DevBuf<unsigned long long> d_sorted;
    int* h_out = nullptr;  // pinned: [corner n_rows*120][ccnt n_rows][pcnt n_rows][poff n_rows+1][planar cap]
    size_t h_cap = 0;
    int smem_sort = 0, smem_ring = 0;
};
FeatWorkspace& workspace(int device) {
    static FeatWorkspace ws[64];
    return ws[device & 63];
}

template <int BLOCK>
void launch_feat(const RingArgs& a, size_t smem_sort, size_t smem_ring, FeatWorkspace& w) {
    if ((int)smem_sort > w.smem_sort || (int)smem_ring > w.smem_ring) {
        FLS_CUDA(cudaFuncSetAttribute(feat_sort_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
        FLS_CUDA(cudaFuncSetAttribute(feat_sort_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
        FLS_CUDA(cudaFuncSetAttribute(feat_ring_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ring));
        FLS_CUDA(cudaFuncSetAttribute(feat_ring_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ring));
        w.smem_sort = (int)smem_sort;
        w.smem_ring = (int)smem_ring;
    }
    feat_sort_kernel<BLOCK><<<dim3(6, a.n_rows), BLOCK, smem_sort, w.st>>>(a);
    feat_ring_kernel<BLOCK><<<a.n_rows, BLOCK, smem_ring, w.st>>>(a);
}

}  // namespace

// Host driver.  Returns fls_status; fills corner_idx / planar_idx (host) in the reference's emission order.
int extract_features_device(int device, const float* depth, const int* col, size_t n, const int* row_start, const int* row_end, int n_rows,
                            float corner_thr, float planar_thr, int* corner_idx, size_t* n_corner, int* planar_idx, size_t* n_planar,
                            fls_match_stats* stats) {
    *n_corner = 0;
    *n_planar = 0;
    if (n < 12 || n_rows <= 0) return FLS_OK;
    if (device < 0 || device >= 64) return FLS_ERR_INVALID_ARG;
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
    const int lcap = max_len + 16;
    const size_t smem_sort = (size_t)lpad * 8;
    const size_t smem_ring = (size_t)lcap * 9 + (size_t)(max_ring > 0 ? max_ring : 1) + 16;
    // block or ring too long for the shared-memory working set (DESIGN.md "limits")
    if (smem_sort > 200 * 1024 || smem_ring > 200 * 1024) return FLS_ERR_UNSUPPORTED;
    FeatWorkspace& w = workspace(device);
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = FLS_OK;
    try {
        FLS_CUDA(cudaSetDevice(device));
        if (!w.ready) {
            FLS_CUDA(cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking));
            FLS_CUDA(cudaEventCreate(&w.e0));
            FLS_CUDA(cudaEventCreate(&w.e1));
            FLS_CUDA(cudaEventCreate(&w.k0));
            FLS_CUDA(cudaEventCreate(&w.k1));
            w.ready = true;
        }
        cudaStream_t st = w.st;
        const size_t o_ccnt = (size_t)n_rows * 120, o_pcnt = o_ccnt + n_rows, o_poff = o_pcnt + n_rows, o_planar = o_poff + n_rows + 1;
        const size_t out_len = o_planar + (size_t)planar_cap + 8;
        w.d_depth.reserve(n);
        w.d_rough.reserve(n);
        w.d_col.reserve(n);
        w.d_meta.reserve(n);
        w.d_sorted.reserve(n);
        w.d_rows.reserve((size_t)n_rows * 2);
        w.d_out.reserve(out_len);
        if (out_len > w.h_cap) {
            if (w.h_out) cudaFreeHost(w.h_out);
            w.h_out = nullptr;
            w.h_cap = 0;
            FLS_CUDA(cudaMallocHost(&w.h_out, out_len * 2 * sizeof(int)));
            w.h_cap = out_len * 2;
        }
        FLS_CUDA(cudaEventRecord(w.e0, st));
        FLS_CUDA(cudaMemcpyAsync(w.d_depth.p, depth, n * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(w.d_col.p, col, n * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(w.d_rows.p, row_start, n_rows * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaMemcpyAsync(w.d_rows.p + n_rows, row_end, n_rows * 4, cudaMemcpyHostToDevice, st));
        FLS_CUDA(cudaEventRecord(w.k0, st));
        feat_point_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.d_depth.p, w.d_col.p, (int)n, w.d_rough.p, w.d_meta.p);
        ring_offsets_kernel<<<1, 32, 0, st>>>(w.d_rows.p, w.d_rows.p + n_rows, n_rows, w.d_out.p + o_poff);
        RingArgs a;
        a.rough = w.d_rough.p;
        a.meta = w.d_meta.p;
        a.row_start = w.d_rows.p;
        a.row_end = w.d_rows.p + n_rows;
        a.n_rows = n_rows;
        a.n = (int)n;
        a.corner_thr = corner_thr;
        a.planar_thr = planar_thr;
        a.lpad = lpad;
        a.lcap = lcap;
        a.ring_cap = max_ring;
        a.max_rounds = 96;
        if (const char* e = std::getenv("FLS_FEAT_MAX_ROUNDS")) a.max_rounds = std::atoi(e) > 0 ? std::atoi(e) : 1;
        a.sorted = w.d_sorted.p;
        a.corner_out = w.d_out.p;
        a.corner_cnt = w.d_out.p + o_ccnt;
        a.planar_cnt = w.d_out.p + o_pcnt;
        a.planar_off = w.d_out.p + o_poff;
        a.planar_out = w.d_out.p + o_planar;
        if (max_len > 512) launch_feat<1024>(a, smem_sort, smem_ring, w);
        else launch_feat<256>(a, smem_sort, smem_ring, w);
        FLS_CUDA(cudaGetLastError());
        FLS_CUDA(cudaEventRecord(w.k1, st));
        FLS_CUDA(cudaMemcpyAsync(w.h_out, w.d_out.p, (o_planar + (size_t)planar_cap) * 4, cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaEventRecord(w.e1, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        const int *h_corner = w.h_out, *h_ccnt = w.h_out + o_ccnt, *h_pcnt = w.h_out + o_pcnt, *h_poff = w.h_out + o_poff,
                  *h_planar = w.h_out + o_planar;
        size_t nc = 0, np = 0;
        for (int r = 0; r < n_rows; ++r) {
            std::memcpy(corner_idx + nc, h_corner + (size_t)r * 120, (size_t)h_ccnt[r] * 4);
            nc += h_ccnt[r];
            std::memcpy(planar_idx + np, h_planar + h_poff[r], (size_t)h_pcnt[r] * 4);
            np += h_pcnt[r];
        }
        *n_corner = nc;
        *n_planar = np;
        if (stats) {
            float ms = 0, kms = 0;
            FLS_CUDA(cudaEventElapsedTime(&ms, w.e0, w.e1));
            FLS_CUDA(cudaEventElapsedTime(&kms, w.k0, w.k1));
            std::memset(stats, 0, sizeof(*stats));
            stats->gpu_ms = ms;
            stats->kernel_ms = kms;
            stats->kernel_launches = 4;
            stats->gpu_launches = 4;
            stats->n_source = (long long)n;
            stats->h2d_bytes = (long long)(n * 8 + (size_t)n_rows * 8);
            stats->d2h_bytes = (long long)((o_planar + (size_t)planar_cap) * 4);
            // algorithmic bytes: depth + col in, roughness/meta/sort keys written and read once, indices out
            stats->algo_bytes = (long long)(n * (4 + 4 + 2 * 4 + 2 * 1 + 2 * 8) + (nc + np) * 4);
        }
    } catch (const CudaError& e) {
        rc = e.status;
    }
    return rc;
}

}  // namespace fls
