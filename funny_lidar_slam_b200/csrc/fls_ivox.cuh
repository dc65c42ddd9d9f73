// fls_ivox.cuh — device-side incremental-voxel map (iVox) and its bounded 5-NN lookup.
//
// Replaces IVoxMap (include/ivox_map/ivox_map.h:16-74, src/ivox_map/ivox_map.cpp upstream):
// the unordered_map + std::list LRU + per-voxel std::vector becomes
//   * one voxel-contiguous SoA float4 point array (points of a voxel adjacent, voxels in Morton order,
//     insertion order preserved inside a voxel), and
//   * an open-addressing table of 16-byte slots {packed key, start, count}, load factor <= 0.5.
#pragma once
#include "fls_common.cuh"

namespace fls {

struct IvoxView {
    const float4* __restrict__ pts;
    const HashSlot* __restrict__ tab;
    unsigned mask;
    float inv_res;
    float max_range2;
    int n_stencil;
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
    __device__ __forceinline__ int count() const { return (j0 != 0xffffffffu) + (j1 != 0xffffffffu) + (j2 != 0xffffffffu) + (j3 != 0xffffffffu) + (j4 != 0xffffffffu); }
};

// IVoxMap::GetClosestPoint(pt, out, 5, max_range) (ivox_map.cpp:6-37 + voxel_grid_node.cpp:23-42 upstream).
// Per-voxel top-K followed by a global top-K equals the global top-K of all in-range candidates, which is
// what is kept here; the nearest ends in slot 0 (the only ordering upstream guarantees).
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

}  // namespace fls
