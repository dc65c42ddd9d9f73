// fls_maps.h — host-side owners of the device-resident map structures.
#pragma once
#include "fls_common.cuh"

namespace fls {

// scratch shared by the sort / run-length passes of every map build
struct BuildScratch {
    DevBuf<unsigned long long> keys, keys_sorted, uniq;
    DevBuf<unsigned> idx, idx_sorted, counts, starts;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<int> num_runs;
    int* h_num_runs = nullptr;  // pinned
    BuildScratch();
    ~BuildScratch();
};

// iVox map: voxel-contiguous float4 points + open-addressing table (see fls_ivox.cuh)
struct IvoxMap {
    float res = 0.5f, inv_res = 2.0f;
    size_t n_pts = 0, n_vox = 0;
    unsigned mask = 0;
    DevBuf<float4> pts_all;     // insertion order (kept so incremental adds can rebuild)
    DevBuf<float4> pts_sorted;  // voxel-contiguous, Morton order
    DevBuf<HashSlot> table;
    BuildScratch scratch;
    int launches = 0;

    void set_resolution(float r) {
        res = r;
        inv_res = 1.0f / r;
    }
    void clear() { n_pts = n_vox = 0; }
    // append n points that are already on the device (packed float4) and rebuild; returns fls_status
    int append_and_build(const float4* d_new, size_t n_new, long long capacity, cudaStream_t st);
    size_t bytes() const { return pts_all.bytes() + pts_sorted.bytes() + table.bytes(); }
};

// repack caller records (stride 16 or >= 20 with intensity at byte 16) into packed float4 on the device
void launch_repack(const unsigned char* d_raw, size_t n, size_t stride, float4* d_out, cudaStream_t st);

}  // namespace fls
