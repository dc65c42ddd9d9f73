// fls_ivox.cuh — device-side incremental-voxel map (iVox) and its bounded 5-NN lookup.
//
// Replaces IVoxMap (include/ivox_map/ivox_map.h:16-74, src/ivox_map/ivox_map.cpp upstream):
// the unordered_map + std::list LRU + per-voxel std::vector becomes
//   * one voxel-contiguous SoA float4 point array (points of a voxel adjacent, voxels in Morton order,
//     insertion order preserved inside a voxel) + an open-addressing table of 16-byte slots
//     {packed key, start, count}, load factor <= 0.5  (the "occupied" table), and
//   * per-centre STENCIL LISTS: for every voxel whose stencil touches an occupied voxel, the candidate
//     sequence GetClosestPoint would visit (stencil order, then insertion order) stored contiguously, with a
//     second table {centre key -> list start, list length}.  One probe + one streaming scan per query.
#pragma once
#include "fls_common.cuh"

namespace fls {

struct IvoxView {
    const float4* __restrict__ pts;    // voxel-contiguous points
    const HashSlot* __restrict__ tab;  // occupied-voxel table
    unsigned mask;
    float inv_res;
    float max_range2;
    int n_stencil;
    const float4* __restrict__ lists;   // stencil lists (may be null)
    const HashSlot* __restrict__ ctab;  // centre table
    unsigned cmask;
    unsigned prefetch;  // 0 none, 1 L2, 2 L1: request the whole candidate run right after the table probe
    unsigned fast_knn;  // 1: every stencil candidate is provably within max_range (range test and exact rounding not needed up front)
};

// stencil offsets in the reference's order (src/ivox_map/ivox_map.cpp:43-66 upstream)
static __device__ __constant__ signed char c_stencil[27][4] = {
    {0, 0, 0, 0},  {-1, 0, 0, 0}, {1, 0, 0, 0},  {0, 1, 0, 0},   {0, -1, 0, 0}, {0, 0, -1, 0},  {0, 0, 1, 0},  {1, 1, 0, 0},  {-1, 1, 0, 0},
    {1, -1, 0, 0}, {-1, -1, 0, 0}, {1, 0, 1, 0}, {-1, 0, 1, 0},  {1, 0, -1, 0}, {-1, 0, -1, 0}, {0, 1, 1, 0},  {0, -1, 1, 0}, {0, 1, -1, 0},
    {0, -1, -1, 0}, {1, 1, 1, 0}, {-1, 1, 1, 0}, {1, -1, 1, 0},  {1, 1, -1, 0}, {-1, -1, 1, 0}, {-1, 1, -1, 0}, {1, -1, -1, 0}, {-1, -1, -1, 0}};

// IVoxMap::Pos2Grid (ivox_map.cpp:145-147 upstream): round(p * inv_res), fp32 product, half away from zero
__device__ __forceinline__ int ivox_coord(float v, float inv_res) { return (int)roundf(__fmul_rn(v, inv_res)); }
// uniform search grid (bounded exact NN): floor(p * inv_cell)
__device__ __forceinline__ int floor_coord(float v, float inv_res) { return (int)floorf(__fmul_rn(v, inv_res)); }
__device__ __forceinline__ int grid_coord(float v, float inv_res, int key_mode) { return key_mode ? floor_coord(v, inv_res) : ivox_coord(v, inv_res); }

__host__ __device__ __forceinline__ unsigned compact21(unsigned long long x) {  // inverse of spread21
    x &= 0x1249249249249249ULL;
    x = (x | x >> 2) & 0x10c30c30c30c30c3ULL;
    x = (x | x >> 4) & 0x100f00f00f00f00fULL;
    x = (x | x >> 8) & 0x1f0000ff0000ffULL;
    x = (x | x >> 16) & 0x1f00000000ffffULL;
    x = (x | x >> 32) & 0x1fffffULL;
    return (unsigned)x;
}
__host__ __device__ __forceinline__ void morton_decode(unsigned long long m, int& x, int& y, int& z) {
    x = (int)compact21(m) - (1 << 20);
    y = (int)compact21(m >> 1) - (1 << 20);
    z = (int)compact21(m >> 2) - (1 << 20);
}

struct Knn5 {
    float d0, d1, d2, d3, d4;
    unsigned j0, j1, j2, j3, j4;
    __device__ __forceinline__ void init() {
        d0 = d1 = d2 = d3 = d4 = INFINITY;
        j0 = j1 = j2 = j3 = j4 = 0xffffffffu;
    }
    // insert keeping ascending (d, visit order): a later candidate never passes an equal earlier one
    __device__ __forceinline__ void push(float d, unsigned j) {
        if (!(d < d4)) return;
        d4 = d;
        j4 = j;
        if (d4 < d3) { float td = d3; d3 = d4; d4 = td; unsigned tj = j3; j3 = j4; j4 = tj; } else return;
        if (d3 < d2) { float td = d2; d2 = d3; d3 = td; unsigned tj = j2; j2 = j3; j3 = tj; } else return;
        if (d2 < d1) { float td = d1; d1 = d2; d2 = td; unsigned tj = j1; j1 = j2; j2 = tj; } else return;
        if (d1 < d0) { float td = d0; d0 = d1; d1 = td; unsigned tj = j0; j0 = j1; j1 = tj; }
    }
};

// IVoxMap::GetClosestPoint(pt, out, 5, max_range) (ivox_map.cpp:6-37 + voxel_grid_node.cpp:23-42 upstream), probing
// the occupied table once per stencil voxel.  Per-voxel top-K followed by a global top-K equals the global top-K of
// all in-range candidates, which is what is kept; the nearest ends in slot 0 (the only ordering upstream guarantees).
// Indices refer to `pts`.
__device__ __forceinline__ void ivox_knn5(const IvoxView& m, float qx, float qy, float qz, Knn5& nn, unsigned& n_cand, unsigned& n_hits) {
    nn.init();
    n_cand = 0;
    n_hits = 0;
    const int kx = ivox_coord(qx, m.inv_res), ky = ivox_coord(qy, m.inv_res), kz = ivox_coord(qz, m.inv_res);
#pragma unroll 1
    for (int s = 0; s < m.n_stencil; ++s) {
        const unsigned long long key = pack_key(kx + c_stencil[s][0], ky + c_stencil[s][1], kz + c_stencil[s][2]);
        unsigned start, count;
        if (!table_find(m.tab, m.mask, key, start, count)) continue;
        n_cand += count;
        n_hits += 1;
#pragma unroll 1
        for (unsigned j = start; j < start + count; ++j) {
            const float4 p = __ldg(m.pts + j);
            const float d = dist2_ref(p.x, p.y, p.z, qx, qy, qz);
            if (d < m.max_range2) nn.push(d, j);
        }
    }
}

// Same result through the stencil lists: one probe of the centre table, then a streaming scan of the
// contiguous candidate run (already in visit order).  Indices refer to `lists`.
__device__ __forceinline__ void ivox_knn5_lists(const IvoxView& m, float qx, float qy, float qz, Knn5& nn, unsigned& n_cand) {
    nn.init();
    n_cand = 0;
    const unsigned long long key = pack_key(ivox_coord(qx, m.inv_res), ivox_coord(qy, m.inv_res), ivox_coord(qz, m.inv_res));
    unsigned start, count;
    if (!table_find(m.ctab, m.cmask, key, start, count)) return;
    n_cand = count;
    const float4* __restrict__ L = m.lists + start;
    unsigned j = 0;
#pragma unroll 1
    for (; j + 4 <= count; j += 4) {
        const float4 p0 = __ldg(L + j), p1 = __ldg(L + j + 1), p2 = __ldg(L + j + 2), p3 = __ldg(L + j + 3);
        const float e0 = dist2_ref(p0.x, p0.y, p0.z, qx, qy, qz);
        const float e1 = dist2_ref(p1.x, p1.y, p1.z, qx, qy, qz);
        const float e2 = dist2_ref(p2.x, p2.y, p2.z, qx, qy, qz);
        const float e3 = dist2_ref(p3.x, p3.y, p3.z, qx, qy, qz);
        if (e0 < m.max_range2) nn.push(e0, start + j);
        if (e1 < m.max_range2) nn.push(e1, start + j + 1);
        if (e2 < m.max_range2) nn.push(e2, start + j + 2);
        if (e3 < m.max_range2) nn.push(e3, start + j + 3);
    }
#pragma unroll 1
    for (; j < count; ++j) {
        const float4 p = __ldg(L + j);
        const float d = dist2_ref(p.x, p.y, p.z, qx, qy, qz);
        if (d < m.max_range2) nn.push(d, start + j);
    }
}

}  // namespace fls
