// fls_knn.cuh — selection pieces of the bounded iVox 5-NN shared by the LOAM-iVox kernels (fls_p2plane.cu, fls_p2plane_v9.cu):
// the exact (fp32 distance, visit order) comparator scan and the quantised top-6 fast path.
#pragma once
#include "fls_common.cuh"
#include "fls_plane.cuh"

namespace fls {
namespace {

// IVoxMap::GetClosestPoint through the stencil lists: one probe of the centre table, then a streaming scan of the
// contiguous candidate run (already in the reference's visit order).  Indices refer to `lists`.
// Exact scan with the full-precision (distance, visit order) comparator — the reference semantics.  Out of line: it is
// the fallback of the quantised fast path below.
__device__ __noinline__ void knn5_exact(const float4* __restrict__ L, unsigned start, unsigned count, float r2, float qx, float qy, float qz,
                                        Top5& nn) {
    nn.init();
    unsigned j = 0;
#pragma unroll 1
    for (; j + 4 <= count; j += 4) {
        const float4 p0 = __ldg(L + j), p1 = __ldg(L + j + 1), p2 = __ldg(L + j + 2), p3 = __ldg(L + j + 3);
        const float e0 = dist2_ref(p0.x, p0.y, p0.z, qx, qy, qz);
        const float e1 = dist2_ref(p1.x, p1.y, p1.z, qx, qy, qz);
        const float e2 = dist2_ref(p2.x, p2.y, p2.z, qx, qy, qz);
        const float e3 = dist2_ref(p3.x, p3.y, p3.z, qx, qy, qz);
        // d < max_range^2 (voxel_grid_node.cpp:27 upstream); rejected candidates become the sentinel key
        const bool i0 = e0 < r2, i1 = e1 < r2, i2 = e2 < r2, i3 = e3 < r2;
        nn.push(i0 ? e0 : INFINITY, i0 ? start + j : 0xffffffffu);
        nn.push(i1 ? e1 : INFINITY, i1 ? start + j + 1 : 0xffffffffu);
        nn.push(i2 ? e2 : INFINITY, i2 ? start + j + 2 : 0xffffffffu);
        nn.push(i3 ? e3 : INFINITY, i3 ? start + j + 3 : 0xffffffffu);
    }
#pragma unroll 1
    for (; j < count; ++j) {
        const float4 p = __ldg(L + j);
        const float d = dist2_ref(p.x, p.y, p.z, qx, qy, qz);
        const bool in = d < r2;
        nn.push(in ? d : INFINITY, in ? start + j : 0xffffffffu);
    }
}

// Fast selection: six 32-bit keys {26 high bits of the fp32 squared distance | 6-bit position in the run}, ordered by
// integer min/max (2 instructions per compare-exchange instead of 5).  Dropping the 6 low mantissa bits can only
// mis-order candidates whose distances agree to 26 bits; that matters for the result only between ranks 1/2 (the
// nearest neighbour anchors the point-to-plane distance) and 5/6 (membership of the 5-NN set) — exactly those two
// pairs are checked afterwards and an ambiguous query (a few per 10 000) is re-run through knn5_exact.
struct Top6q {
    unsigned k0, k1, k2, k3, k4, k5;
    __device__ __forceinline__ void init() { k0 = k1 = k2 = k3 = k4 = k5 = 0xffffffffu; }
#define FLS_CEQ(a, b)                 \
    {                                 \
        const unsigned lo_ = min(a, b); \
        b = max(a, b);                \
        a = lo_;                      \
    }
    __device__ __forceinline__ void push(unsigned key) {
        k5 = min(k5, key);
        FLS_CEQ(k4, k5)
        FLS_CEQ(k3, k4)
        FLS_CEQ(k2, k3)
        FLS_CEQ(k1, k2)
        FLS_CEQ(k0, k1)
    }
#undef FLS_CEQ
};

__device__ __forceinline__ float dist2_fast(float px, float py, float pz, float qx, float qy, float qz) {
    const float dx = px - qx, dy = py - qy, dz = pz - qz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
__device__ __forceinline__ unsigned qkey(float d, float r2, unsigned jl) {
    return (d < r2) ? ((__float_as_uint(d) & 0xffffffc0u) | jl) : 0xffffffffu;  // d < max_range^2 (voxel_grid_node.cpp:27 upstream)
}

}  // namespace
}  // namespace fls
