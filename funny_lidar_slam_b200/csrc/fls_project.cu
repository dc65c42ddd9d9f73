// fls_project.cu — PointcloudProjector::Project (src/loam/pointcloud_projector.cpp:32-133 upstream) on the device:
// range gate, column from the 7th-order FastAtan2 polynomial (include/common/math_function.h:159-186), "first hit wins"
// per (ring, column) cell, then the row-major compaction that produces ordered_cloud_ / point_depth_vec_ /
// point_col_index_vec_ / row_start_index_vec_ / row_end_index_vec_ — the arrays the feature extractor (K4) consumes.
//
// The sequential "first point to reach a cell keeps it" (:90-91) becomes an atomicMin on the point's position in the
// raw cloud: the smallest index is by definition the first one the sequential loop would have seen.  The per-point
// de-skew (ProcessPoint, :100-103) runs on the winner of every cell when the caller passes the IMU orientation buffer
// (fls_project_imu; fls_deskew.cuh); without one the points pass through unchanged.
#include <cub/cub.cuh>

#include <mutex>

#include "fls_deskew.cuh"
#include "fls_maps.h"

namespace fls {
namespace {

__device__ __forceinline__ float fast_atan2_ref(float y, float x) {
    const float p1 = 0.9997878412794807f, p3 = -0.3258083974640975f, p5 = 0.1555786518463281f, p7 = -0.04432655554792128f;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = 1.1920928955078125e-07f;
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        const float c2 = __fmul_rn(c, c);
        a = __fsub_rn(1.57079632679489661923f,
                      __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(3.14159265358979323846f, a);
    if (y < 0.f) a = __fsub_rn(6.28318530717958647692f, a);
    if (a > 3.14159265358979323846f) a = __fsub_rn(a, 6.28318530717958647692f);
    return a;
}

__device__ __forceinline__ float depth_ref(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void proj_clear_kernel(unsigned* __restrict__ winner, size_t cells) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < cells) winner[i] = 0xffffffffu;
}

__global__ void proj_claim_kernel(const float4* __restrict__ raw, const int* __restrict__ ring, const float* __restrict__ time, DeskewView dv, int n,
                                  int V, int H, float h_res, float min_d, float max_d, unsigned* __restrict__ winner) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = raw[k];
    const float depth = depth_ref(p.x, p.y, p.z);
    if (depth < min_d || depth > max_d) return;  // :59-61
    const int row = ring[k];
    int col = (int)roundf(__fdiv_rn(fast_atan2_ref(p.y, p.x), h_res)) + H / 2;  // :64-65
    if (col >= H) col -= H;
    if (row >= V || row < 0 || col < 0 || col >= H) return;  // :81-82
    if (dv.m > 0) {  // ProcessPoint fails for a time outside the IMU buffer: the point returns without claiming its cell (:100-103)
        const unsigned long long t = (unsigned long long)((long long)dv.ref_time + (long long)__dmul_rn((double)time[k], 1.0e6));
        if (dv.m < 2 || dv.t[0] > t || dv.t[dv.m - 1] < t) return;
    }
    atomicMin(&winner[(size_t)row * H + col], (unsigned)k);  // :86-87 first hit wins
}

__global__ void proj_flags_kernel(const unsigned* __restrict__ winner, size_t cells, unsigned* __restrict__ flag) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < cells) flag[i] = winner[i] != 0xffffffffu ? 1u : 0u;
}

__global__ void proj_emit_kernel(const float4* __restrict__ raw, const float* __restrict__ time, DeskewView dv, const unsigned* __restrict__ winner,
                                 const unsigned* __restrict__ excl, int V, int H,
                                 float4* __restrict__ ordered, float* __restrict__ depth, int* __restrict__ col, int* __restrict__ row_start,
                                 int* __restrict__ row_end, unsigned* __restrict__ total) {
    const size_t cells = (size_t)V * H;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const unsigned w = winner[i];
    const unsigned pos = excl[i];
    if (w != 0xffffffffu) {
        const float4 p = raw[w];
        float4 q = p;
        if (dv.m > 0) deskew_point(dv, p.x, p.y, p.z, time[w], q.x, q.y, q.z);  // :100-103; the range stays the raw point's (:57, :105)
        ordered[pos] = q;
        depth[pos] = depth_ref(p.x, p.y, p.z);
        col[pos] = (int)(i % (size_t)H);
    }
    const int c = (int)(i % (size_t)H), r = (int)(i / (size_t)H);
    if (c == 0) row_start[r] = (int)pos + 5;  // :115
    if (c == H - 1) {
        const unsigned end = pos + (w != 0xffffffffu ? 1u : 0u);
        row_end[r] = (int)end - 6;  // :131
        if (r == V - 1) *total = end;
    }
}

struct ProjWorkspace {
    std::mutex mu;
    bool ready = false;
    cudaStream_t st = nullptr;
    DevBuf<float4> raw, ordered;
    DevBuf<unsigned char> staging;
    DevBuf<int> ring, col, rows;
    DevBuf<unsigned> winner, flag, excl, total;
    DevBuf<float> depth, time;
    DevBuf<unsigned long long> imu_t;
    DevBuf<double> imu_q;
    DevBuf<unsigned char> cub_tmp;
};
ProjWorkspace& proj_workspace(int device) {
    static ProjWorkspace ws[64];
    return ws[device & 63];
}

}  // namespace

// Host driver: raw cloud (host, `stride` bytes per record) + ring per point -> projector arrays (host).  depth_out / col_out hold V*H
// entries (the first *n_out are meaningful, as upstream), ordered_out V*H packed float4 records.
int make_deskew_view(const fls_imu_buffer* imu, DevBuf<unsigned long long>& d_t, DevBuf<double>& d_q, cudaStream_t st, DeskewView& v);

int project_device(int device, const void* raw, const int* ring, const float* time, const fls_imu_buffer* imu, size_t n, size_t stride, int V, int H,
                   float h_res, float min_d, float max_d, float* ordered_out, float* depth_out, int* col_out, int* row_start, int* row_end, size_t* n_out) {
    *n_out = 0;
    if (V <= 0 || H <= 0 || !(h_res > 0.f) || device < 0 || device >= 64 || n > 0x7fffffffull) return FLS_ERR_INVALID_ARG;
    const size_t cells = (size_t)V * H;
    if (cells > 0x7fffffffull) return FLS_ERR_INVALID_ARG;
    ProjWorkspace& w = proj_workspace(device);
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = FLS_OK;
    try {
        FLS_CUDA(cudaSetDevice(device));
        if (!w.ready) {
            FLS_CUDA(cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking));
            w.ready = true;
        }
        cudaStream_t st = w.st;
        DeskewView dv;
        const bool use_imu = imu && imu->n_imu && time;
        rc = make_deskew_view(use_imu ? imu : nullptr, w.imu_t, w.imu_q, st, dv);
        const bool ref_failed = rc == FLS_ERR_INVALID_ARG && use_imu && imu->imu_time_us && imu->imu_quat_xyzw;  // SetRefTime failed: no point is accepted
        if (rc != FLS_OK && !ref_failed) return rc;
        rc = FLS_OK;
        if (ref_failed) n = 0;
        w.raw.reserve(n + 1);
        w.ring.reserve(n + 1);
        w.time.reserve(n + 1);
        if (n && use_imu) FLS_CUDA(cudaMemcpyAsync(w.time.p, time, n * sizeof(float), cudaMemcpyHostToDevice, st));
        w.winner.reserve(cells);
        w.flag.reserve(cells);
        w.excl.reserve(cells);
        w.total.reserve(1);
        w.ordered.reserve(cells);
        w.depth.reserve(cells);
        w.col.reserve(cells);
        w.rows.reserve((size_t)V * 2);
        if (n) {
            if (stride == FLS_LAYOUT_PACKED) {
                FLS_CUDA(cudaMemcpyAsync(w.raw.p, raw, n * sizeof(float4), cudaMemcpyHostToDevice, st));
            } else {
                w.staging.reserve(n * stride);
                FLS_CUDA(cudaMemcpyAsync(w.staging.p, raw, n * stride, cudaMemcpyHostToDevice, st));
                launch_repack(w.staging.p, n, stride, w.raw.p, st);
            }
            FLS_CUDA(cudaMemcpyAsync(w.ring.p, ring, n * sizeof(int), cudaMemcpyHostToDevice, st));
        }
        const unsigned gc = (unsigned)((cells + 255) / 256);
        proj_clear_kernel<<<gc, 256, 0, st>>>(w.winner.p, cells);
        if (n) proj_claim_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.raw.p, w.ring.p, w.time.p, dv, (int)n, V, H, h_res, min_d, max_d, w.winner.p);
        proj_flags_kernel<<<gc, 256, 0, st>>>(w.winner.p, cells, w.flag.p);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, w.flag.p, w.excl.p, (int)cells, st);
        w.cub_tmp.reserve(tb + 256);
        tb = w.cub_tmp.cap;
        FLS_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp.p, tb, w.flag.p, w.excl.p, (int)cells, st));
        // upstream leaves the tails of depth / col as they were (resize to V*H, col zero-filled): zero both
        FLS_CUDA(cudaMemsetAsync(w.depth.p, 0, cells * sizeof(float), st));
        FLS_CUDA(cudaMemsetAsync(w.col.p, 0, cells * sizeof(int), st));
        proj_emit_kernel<<<gc, 256, 0, st>>>(w.raw.p, w.time.p, dv, w.winner.p, w.excl.p, V, H, w.ordered.p, w.depth.p, w.col.p, w.rows.p, w.rows.p + V, w.total.p);
        FLS_CUDA(cudaGetLastError());
        unsigned total = 0;
        FLS_CUDA(cudaMemcpyAsync(&total, w.total.p, sizeof(total), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(depth_out, w.depth.p, cells * sizeof(float), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(col_out, w.col.p, cells * sizeof(int), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(row_start, w.rows.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(row_end, w.rows.p + V, (size_t)V * sizeof(int), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        if (total) FLS_CUDA(cudaMemcpy(ordered_out, w.ordered.p, (size_t)total * sizeof(float4), cudaMemcpyDeviceToHost));
        *n_out = total;
    } catch (const CudaError& e) {
        rc = e.status;
    }
    return rc;
}

}  // namespace fls
