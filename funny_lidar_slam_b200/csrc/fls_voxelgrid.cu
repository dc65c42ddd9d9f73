// fls_voxelgrid.cu — K7: pcl::VoxelGrid<PointXYZI>::filter as wrapped by VoxelGridCloud
// (include/common/pointcloud_utility.h:216-224,263-271 upstream; PCL 1.10 semantics):
//   bounding box -> cell = floor(x*inv_leaf) - min_b (fp32) -> linear id -> sort by id -> one fp32 centroid
//   (xyz AND intensity) per occupied cell, cells in ascending id; dx*dy*dz > INT_MAX returns the input unchanged.
// The radix sort is stable, so every centroid is summed in input order — the order the oracle pins.
#include <cub/cub.cuh>

#include "fls_maps.h"

namespace fls {
namespace {

struct MinMax6 {
    float mn[3], mx[3];
};

// order-preserving float <-> uint encoding so the bounding box can be reduced with integer atomicMin / atomicMax
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned o) {
    const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

struct MinMaxOrd {
    unsigned mn[3], mx[3];
};

__global__ void minmax_kernel(const float4* __restrict__ pts, size_t n, MinMaxOrd* __restrict__ out) {
    __shared__ float s[6][256];
    float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        mn0 = fminf(mn0, p.x); mx0 = fmaxf(mx0, p.x);
        mn1 = fminf(mn1, p.y); mx1 = fmaxf(mx1, p.y);
        mn2 = fminf(mn2, p.z); mx2 = fmaxf(mx2, p.z);
    }
    s[0][threadIdx.x] = mn0; s[1][threadIdx.x] = mn1; s[2][threadIdx.x] = mn2;
    s[3][threadIdx.x] = mx0; s[4][threadIdx.x] = mx1; s[5][threadIdx.x] = mx2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            for (int k = 0; k < 3; ++k) s[k][threadIdx.x] = fminf(s[k][threadIdx.x], s[k][threadIdx.x + o]);
            for (int k = 3; k < 6; ++k) s[k][threadIdx.x] = fmaxf(s[k][threadIdx.x], s[k][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) atomicMin(&out->mn[threadIdx.x], f2ord(s[threadIdx.x][0]));
    else if (threadIdx.x < 6) atomicMax(&out->mx[threadIdx.x - 3], f2ord(s[threadIdx.x][0]));
}

__global__ void minmax_init_kernel(MinMaxOrd* o) {
    for (int k = 0; k < 3; ++k) {
        o->mn[k] = 0xffffffffu;
        o->mx[k] = 0u;
    }
}

__global__ void vg_keys_kernel(const float4* __restrict__ pts, size_t n, float inv, int mb0, int mb1, int mb2, int mul1, int mul2,
                               unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const int i0 = (int)(floorf(__fmul_rn(p.x, inv)) - (float)mb0);
    const int i1 = (int)(floorf(__fmul_rn(p.y, inv)) - (float)mb1);
    const int i2 = (int)(floorf(__fmul_rn(p.z, inv)) - (float)mb2);
    keys[i] = (unsigned)(i0 + i1 * mul1 + i2 * mul2);
    idx[i] = (unsigned)i;
}

// one thread per occupied cell: sequential fp32 sums in input order (CentroidPoint accumulators of PCL)
__global__ void vg_centroid_kernel(const float4* __restrict__ pts, const unsigned* __restrict__ idx_sorted, const unsigned* __restrict__ starts,
                                   const unsigned* __restrict__ counts, int runs, float4* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= runs) return;
    const unsigned s = starts[r], c = counts[r];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (unsigned k = 0; k < c; ++k) {
        const float4 p = __ldg(pts + idx_sorted[s + k]);
        sx = __fadd_rn(sx, p.x);
        sy = __fadd_rn(sy, p.y);
        sz = __fadd_rn(sz, p.z);
        si = __fadd_rn(si, p.w);
    }
    const float n = (float)c;
    out[r] = make_float4(__fdiv_rn(sx, n), __fdiv_rn(sy, n), __fdiv_rn(sz, n), __fdiv_rn(si, n));
}

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

// Returns the number of output points; d_out must hold n records.  Synchronises the stream twice
// (bounding box, run count) — both values size the following launches.
size_t voxel_grid_device(const float4* d_pts, size_t n, float leaf, float4* d_out, BuildScratch& sc, cudaStream_t st, int* launches) {
    if (n == 0) return 0;
    const float inv = 1.0f / leaf;
    sc.minmax.reserve(16);
    MinMaxOrd* d_mm = reinterpret_cast<MinMaxOrd*>(sc.minmax.p);
    sc.num_runs.reserve(2);
    minmax_init_kernel<<<1, 1, 0, st>>>(d_mm);
    const unsigned g = grid_for(n, 256) < 592 ? grid_for(n, 256) : 592;
    minmax_kernel<<<g, 256, 0, st>>>(d_pts, n, d_mm);
    MinMaxOrd ho;
    FLS_CUDA(cudaMemcpyAsync(&ho, d_mm, sizeof(ho), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    MinMax6 h;
    for (int a = 0; a < 3; ++a) {
        h.mn[a] = ord2f(ho.mn[a]);
        h.mx[a] = ord2f(ho.mx[a]);
    }
    if (launches) *launches += 2;
    const long long dx = (long long)((h.mx[0] - h.mn[0]) * inv) + 1, dy = (long long)((h.mx[1] - h.mn[1]) * inv) + 1,
                    dz = (long long)((h.mx[2] - h.mn[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647LL) {  // PCL: "Leaf size is too small" -> output = input
        FLS_CUDA(cudaMemcpyAsync(d_out, d_pts, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        return n;
    }
    int minb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
        minb[a] = (int)floorf(h.mn[a] * inv);
        const int maxb = (int)floorf(h.mx[a] * inv);
        divb[a] = maxb - minb[a] + 1;
    }
    const int mul1 = divb[0], mul2 = divb[0] * divb[1];
    sc.idx.reserve(n);
    sc.idx_sorted.reserve(n);
    sc.counts.reserve(n);
    sc.starts.reserve(n);
    sc.k32a.reserve(n);
    sc.k32b.reserve(n);
    sc.uniq32.reserve(n);
    vg_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_pts, n, inv, minb[0], minb[1], minb[2], mul1, mul2, sc.k32a.p, sc.idx.p);
    int end_bit = 1;
    {
        const long long maxid = (long long)divb[0] * divb[1] * divb[2];
        while ((1LL << end_bit) < maxid && end_bit < 32) ++end_bit;
    }
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, sc.k32a.p, sc.k32b.p, sc.idx.p, sc.idx_sorted.p, (int)n, 0, end_bit, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, t2, sc.k32b.p, sc.uniq32.p, sc.counts.p, sc.num_runs.p, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, sc.counts.p, sc.starts.p, (int)n, st);
    size_t tmp = t1 > t2 ? t1 : t2;
    tmp = tmp > t3 ? tmp : t3;
    sc.cub_tmp.reserve(tmp + 256);
    size_t tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortPairs(sc.cub_tmp.p, tb, sc.k32a.p, sc.k32b.p, sc.idx.p, sc.idx_sorted.p, (int)n, 0, end_bit, st));
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRunLengthEncode::Encode(sc.cub_tmp.p, tb, sc.k32b.p, sc.uniq32.p, sc.counts.p, sc.num_runs.p, (int)n, st));
    FLS_CUDA(cudaMemcpyAsync(sc.h_num_runs, sc.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const int runs = *sc.h_num_runs;
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceScan::ExclusiveSum(sc.cub_tmp.p, tb, sc.counts.p, sc.starts.p, runs, st));
    vg_centroid_kernel<<<grid_for(runs, 128), 128, 0, st>>>(d_pts, sc.idx_sorted.p, sc.starts.p, sc.counts.p, runs, d_out);
    FLS_CUDA(cudaGetLastError());
    if (launches) *launches += 6;
    return (size_t)runs;
}

}  // namespace fls
