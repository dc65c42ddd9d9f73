// fls_handle.h — the object behind `fls_handle*`: configuration, stream, device-resident map and scan state.
#pragma once
#include <vector>

#include "fls_common.cuh"
#include "fls_kernels.h"
#include "fls_maps.h"

namespace fls {

struct Handle {
    fls_config cfg;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // per-call accounting (fls_match_stats)
    int launches = 0;
    long long h2d_bytes = 0, d2h_bytes = 0;
    float last_gpu_ms = 0.f;

    // scan-side buffers
    DevBuf<unsigned char> raw;  // strided caller records before repacking
    DevBuf<float4> src;         // scan entering the GN loop (packed float4)
    DevBuf<float4> stage;       // clouds handed to AddCloudToLocalMap
    DevBuf<float4> rec0, rec1;  // persistent per-point {J, |d|} records (LOAM plug-ins)
    DevBuf<unsigned char> flags;
    DevBuf<double> partials;
    DevBuf<GnState> state;
    GnState* h_state = nullptr;  // pinned
    DevBuf<fls_iter_log> log;
    std::vector<fls_iter_log> h_log;
    int log_cap = 0, log_n = 0;
    double T_final[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<cudaEvent_t> prof_ev;  // FLS_FLAG_PROFILE: 2 events per iteration around the residual kernel
    bool profile = false;
    long long per_point_iter_bytes = 0;  // fixed part of the algorithmic bytes per point-iteration (set by match_*)
    long long per_cand_bytes = 0;        // bytes per scanned map record

    // maps
    IvoxMap ivox;

    explicit Handle(const fls_config& c);
    ~Handle();
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;

    void begin_call();
    void end_call(fls_match_stats* st);
    const float4* upload(const void* pts, size_t n, size_t stride, DevBuf<float4>& dst);
    IvoxView ivox_view() const;
    int finish_match(double* T, int* converged, fls_match_stats* st, long long n_source);

    int add_cloud_ivox(const void* pts, size_t n, size_t stride);
    int match_p2plane_ivox(const float4* d_src, size_t n, double* T, int* converged, fls_match_stats* st);
};

}  // namespace fls
