// fls_preprocess.cu — the point pipeline in front of the registration plug-ins, on the device:
// PreProcessing::Run's branch for PointToPlane_IVOX / PointToPlane_KdTree / IcpOptimized / IncrementalNDT
// (src/slam/preprocessing.cpp:181-225 upstream): range gate -> IMU de-skew of every point (fls_deskew.cuh) -> ordered_cloud_
// (every kept point) and planar_cloud_ (every lidar_point_jump_span-th RAW index among the kept ones, then pcl::VoxelGrid).
#include <cub/cub.cuh>

#include <cmath>
#include <mutex>

#include "fls_deskew.cuh"
#include "fls_maps.h"

namespace fls {

bool deskew_ref_inverse(const unsigned long long* t, const double* q, size_t m, unsigned long long ref, double* qri) {
    if (m < 2 || t[0] > ref || t[m - 1] < ref) return false;
    size_t l;
    if (t[0] == ref) l = 0;
    else if (t[m - 1] == ref) l = m - 2;
    else {
        l = m - 1;
        while (ref < t[l]) --l;
    }
    const size_t r = l + 1;
    const double s = double(ref - t[l]) / double(t[r] - t[l]);
    const double u = 1.0 - s;
    volatile double qx = q[4 * l] * u + q[4 * r] * s, qy = q[4 * l + 1] * u + q[4 * r + 1] * s, qz = q[4 * l + 2] * u + q[4 * r + 2] * s,
                    qw = q[4 * l + 3] * u + q[4 * r + 3] * s;
    const double n = std::sqrt(((qx * qx + qy * qy) + qz * qz) + qw * qw);
    const double nx = qx / n, ny = qy / n, nz = qz / n, nw = qw / n;
    const double n2 = ((nx * nx + ny * ny) + nz * nz) + nw * nw;
    qri[0] = -nx / n2;
    qri[1] = -ny / n2;
    qri[2] = -nz / n2;
    qri[3] = nw / n2;
    return true;
}

namespace {

__device__ __forceinline__ float depth_ref_pp(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// one thread per raw point: gate + de-skew; flags: bit 0 kept (ordered cloud), bit 1 kept and i % jump == 0 (planar candidates)
__global__ void pp_point_kernel(const float* __restrict__ raw /*x,y,z,i,t*/, int n, DeskewView dv, float min_d, float max_d, int jump,
                                float4* __restrict__ corrected, unsigned* __restrict__ f_ord, unsigned* __restrict__ f_pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = raw + 5 * (size_t)i;
    float x = p[0], y = p[1], z = p[2];
    bool keep = true;
    const float depth = depth_ref_pp(x, y, z);
    if (depth < min_d || depth > max_d) keep = false;  // preprocessing.cpp:199-203
    if (keep && dv.m > 0) keep = deskew_point(dv, x, y, z, p[4], x, y, z);  // :205-211
    corrected[i] = make_float4(x, y, z, p[3]);
    f_ord[i] = keep ? 1u : 0u;
    f_pl[i] = (keep && (i % jump) == 0) ? 1u : 0u;  // :219-222
}
__global__ void pp_scatter_kernel(const float4* __restrict__ corrected, int n, const unsigned* __restrict__ f, const unsigned* __restrict__ excl,
                                  float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && f[i]) out[excl[i]] = corrected[i];
}

struct PpWorkspace {
    std::mutex mu;
    bool ready = false;
    cudaStream_t st = nullptr;
    DevBuf<float> raw;
    DevBuf<float4> corrected, ordered, planar, filtered;
    DevBuf<unsigned> f_ord, f_pl, e_ord, e_pl;
    DevBuf<unsigned long long> imu_t;
    DevBuf<double> imu_q;
    DevBuf<unsigned char> cub_tmp;
    BuildScratch scratch;
};
PpWorkspace& pp_workspace(int device) {
    static PpWorkspace ws[64];
    return ws[device & 63];
}

}  // namespace

// uploads the IMU samples and fills the view; returns FLS_ERR_INVALID_ARG when ref_time is outside the buffer (SetRefTime fails)
int make_deskew_view(const fls_imu_buffer* imu, DevBuf<unsigned long long>& d_t, DevBuf<double>& d_q, cudaStream_t st, DeskewView& v) {
    std::memset(&v, 0, sizeof(v));
    if (!imu || imu->n_imu == 0) return FLS_OK;  // identity
    if (!imu->imu_time_us || !imu->imu_quat_xyzw || imu->n_imu > 0x7fffffffull) return FLS_ERR_INVALID_ARG;
    if (!deskew_ref_inverse(reinterpret_cast<const unsigned long long*>(imu->imu_time_us), imu->imu_quat_xyzw, imu->n_imu, imu->ref_time_us, v.qri))
        return FLS_ERR_INVALID_ARG;
    d_t.reserve(imu->n_imu);
    d_q.reserve(imu->n_imu * 4);
    FLS_CUDA(cudaMemcpyAsync(d_t.p, imu->imu_time_us, imu->n_imu * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    FLS_CUDA(cudaMemcpyAsync(d_q.p, imu->imu_quat_xyzw, imu->n_imu * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    v.t = d_t.p;
    v.q = d_q.p;
    v.m = (int)imu->n_imu;
    v.ref_time = imu->ref_time_us;
    for (int k = 0; k < 16; ++k) v.T[k] = imu->T_lidar_to_imu[k];
    return FLS_OK;
}

int preprocess_device(int device, const float* raw_xyzit, size_t n, const fls_imu_buffer* imu, float min_d, float max_d, int jump_span, float leaf,
                      float* ordered_out, size_t* n_ordered, float* planar_out, size_t* n_planar) {
    *n_ordered = *n_planar = 0;
    if (device < 0 || device >= 64 || n > 0x7fffffffull || jump_span < 1 || !(leaf > 0.f)) return FLS_ERR_INVALID_ARG;
    if (n == 0) return FLS_OK;
    PpWorkspace& w = pp_workspace(device);
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
        rc = make_deskew_view(imu, w.imu_t, w.imu_q, st, dv);
        if (rc != FLS_OK) return rc == FLS_ERR_INVALID_ARG && imu && imu->n_imu ? FLS_OK : rc;  // SetRefTime failed: upstream drops the scan (:178-183) -> empty clouds
        w.raw.reserve(n * 5);
        w.corrected.reserve(n);
        w.ordered.reserve(n);
        w.planar.reserve(n);
        w.filtered.reserve(n);
        w.f_ord.reserve(n);
        w.f_pl.reserve(n);
        w.e_ord.reserve(n);
        w.e_pl.reserve(n);
        FLS_CUDA(cudaMemcpyAsync(w.raw.p, raw_xyzit, n * 5 * sizeof(float), cudaMemcpyHostToDevice, st));
        const unsigned g = (unsigned)((n + 255) / 256);
        pp_point_kernel<<<g, 256, 0, st>>>(w.raw.p, (int)n, dv, min_d, max_d, jump_span, w.corrected.p, w.f_ord.p, w.f_pl.p);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, w.f_ord.p, w.e_ord.p, (int)n, st);
        w.cub_tmp.reserve(tb + 256);
        tb = w.cub_tmp.cap;
        FLS_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp.p, tb, w.f_ord.p, w.e_ord.p, (int)n, st));
        tb = w.cub_tmp.cap;
        FLS_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp.p, tb, w.f_pl.p, w.e_pl.p, (int)n, st));
        pp_scatter_kernel<<<g, 256, 0, st>>>(w.corrected.p, (int)n, w.f_ord.p, w.e_ord.p, w.ordered.p);
        pp_scatter_kernel<<<g, 256, 0, st>>>(w.corrected.p, (int)n, w.f_pl.p, w.e_pl.p, w.planar.p);
        unsigned last[4] = {0, 0, 0, 0};
        FLS_CUDA(cudaMemcpyAsync(&last[0], w.e_ord.p + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(&last[1], w.f_ord.p + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(&last[2], w.e_pl.p + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaMemcpyAsync(&last[3], w.f_pl.p + (n - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        const size_t no = (size_t)last[0] + last[1], npl = (size_t)last[2] + last[3];
        int l = 0;
        const size_t nf = npl ? voxel_grid_device(w.planar.p, npl, leaf, w.filtered.p, w.scratch, st, &l) : 0;  // :224-225
        if (no) FLS_CUDA(cudaMemcpyAsync(ordered_out, w.ordered.p, no * sizeof(float4), cudaMemcpyDeviceToHost, st));
        if (nf) FLS_CUDA(cudaMemcpyAsync(planar_out, w.filtered.p, nf * sizeof(float4), cudaMemcpyDeviceToHost, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        *n_ordered = no;
        *n_planar = nf;
    } catch (const CudaError& e) {
        rc = e.status;
    }
    return rc;
}

}  // namespace fls
