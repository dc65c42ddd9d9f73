// fls_api.cu — the C ABI of include/fls_b200.h: handle lifetime, host<->device staging and the host side of
// each plug-in's Match / AddCloudToLocalMap / GetFitnessScore.  The Gauss-Newton loop itself runs on the
// device (residual kernel + gn_solve kernel per iteration, convergence decided on the device); the host
// enqueues the iteration cap and reads the 1 KB state block back once.
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "fls_gn.cuh"
#include "fls_handle.h"

namespace fls {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }

Handle::Handle(const fls_config& c) : cfg(c) {
    FLS_CUDA(cudaSetDevice(cfg.device));
    FLS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FLS_CUDA(cudaEventCreate(&ev0));
    FLS_CUDA(cudaEventCreate(&ev1));
    FLS_CUDA(cudaMallocHost(&h_state, sizeof(GnState)));
    state.reserve(1);
    ivox.set_resolution(cfg.ivox_resolution);
    profile = (cfg.flags & FLS_FLAG_PROFILE) != 0;
    if (profile) {
        prof_ev.resize(2 * (size_t)(cfg.max_iterations > 0 ? cfg.max_iterations : 1));
        for (auto& e : prof_ev) FLS_CUDA(cudaEventCreate(&e));
    }
    if (cfg.flags & FLS_FLAG_ITER_LOG) {
        log_cap = cfg.max_iterations > 0 ? cfg.max_iterations : 1;
        log.reserve(log_cap);
        h_log.resize(log_cap);
    }
}

Handle::~Handle() {
    cudaSetDevice(cfg.device);
    if (stream) cudaStreamSynchronize(stream);
    if (h_state) cudaFreeHost(h_state);
    for (auto& e : prof_ev) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

// Copy a caller cloud (host memory, `stride` bytes per record) into a packed float4 device buffer.
const float4* Handle::upload(const void* pts, size_t n, size_t stride, DevBuf<float4>& dst) {
    dst.reserve(n);
    if (n == 0) return dst.p;
    if (stride == FLS_LAYOUT_PACKED) {
        FLS_CUDA(cudaMemcpyAsync(dst.p, pts, n * 16, cudaMemcpyHostToDevice, stream));
        h2d_bytes += (long long)(n * 16);
    } else {
        raw.reserve(n * stride);
        FLS_CUDA(cudaMemcpyAsync(raw.p, pts, n * stride, cudaMemcpyHostToDevice, stream));
        h2d_bytes += (long long)(n * stride);
        launch_repack(raw.p, n, stride, dst.p, stream);
        launches++;
    }
    return dst.p;
}

IvoxView Handle::ivox_view() const {
    IvoxView v;
    v.pts = ivox.pts_sorted.p;
    v.tab = ivox.table.p;
    v.mask = ivox.mask;
    v.inv_res = ivox.inv_res;
    v.max_range2 = cfg.ivox_max_range * cfg.ivox_max_range;
    static const int counts[4] = {1, 7, 19, 27};
    v.n_stencil = counts[cfg.ivox_nearby];
    return v;
}

void Handle::begin_call() {
    FLS_CUDA(cudaSetDevice(cfg.device));
    launches = 0;
    h2d_bytes = d2h_bytes = 0;
    FLS_CUDA(cudaEventRecord(ev0, stream));
}

void Handle::end_call(fls_match_stats* st) {
    FLS_CUDA(cudaEventRecord(ev1, stream));
    FLS_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    FLS_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_gpu_ms = ms;
    if (st) {
        st->gpu_ms = ms;
        st->gpu_launches = launches;
        st->h2d_bytes = h2d_bytes;
        st->d2h_bytes = d2h_bytes;
    }
}

// ---- LoamPointToPlaneIVOX ------------------------------------------------------------------------------------
int Handle::add_cloud_ivox(const void* pts, size_t n, size_t stride) {
    if (cfg.localization_mode) ivox.clear();  // loam_point_to_plane_ivox.h:64-69 upstream: map re-created per call
    else if (ivox.n_pts != 0) return FLS_ERR_UNSUPPORTED;  // external non-first insert relies on Match-internal caches upstream
    const float4* d = upload(pts, n, stride, stage);
    const int rc = ivox.append_and_build(d, n, cfg.ivox_capacity, stream);
    launches += ivox.launches;
    ivox.launches = 0;
    return rc;
}

int Handle::match_p2plane_ivox(const float4* d_src, size_t n, double* T, int* converged, fls_match_stats* st) {
    if (ivox.n_pts == 0) return FLS_ERR_NO_MAP;
    const int ni = (int)n;
    const int grid = p2plane_grid(ni);
    rec0.reserve(n);
    rec1.reserve(n);
    flags.reserve(n);
    partials.reserve((size_t)(grid > 0 ? grid : 1) * kAccStride);
    if (n) FLS_CUDA(cudaMemsetAsync(flags.p, 0, n, stream));
    launch_gn_init(state.p, T, stream);
    launches++;
    P2PlaneArgs a;
    a.src = d_src;
    a.n = ni;
    a.map = ivox_view();
    a.plane_thres = cfg.point_to_planar_thres;
    a.state = state.p;
    a.rec0 = rec0.p;
    a.rec1 = rec1.p;
    a.flags = flags.p;
    a.partials = partials.p;
    GnParams gp;
    gp.method = FLS_P2PLANE_IVOX;
    gp.max_iterations = cfg.max_iterations;
    gp.min_effective = 50;
    gp.n_blocks = grid;
    gp.rot_thres = cfg.rotation_converge_thres;
    gp.pos_thres = cfg.position_converge_thres;
    // roofline accounting: 16 B source point + n_stencil x 16 B slot probes + 32 B persistent
    // record per point-iteration, 16 B per map record resident in the hit voxels.
    per_point_iter_bytes = 16 + 16LL * a.map.n_stencil + 32;
    per_cand_bytes = 16;
    for (int it = 0; it < cfg.max_iterations; ++it) {
        if (profile) FLS_CUDA(cudaEventRecord(prof_ev[2 * it], stream));
        launch_p2plane_iter(a, stream);
        if (profile) FLS_CUDA(cudaEventRecord(prof_ev[2 * it + 1], stream));
        launch_gn_solve(state.p, partials.p, gp, log.p, log_cap, stream);
        launches += (ni > 0 ? 2 : 1);
    }
    return finish_match(T, converged, st, (long long)n);
}

// read the device state back, fill T / stats / iteration log
int Handle::finish_match(double* T, int* converged, fls_match_stats* st, long long n_source) {
    FLS_CUDA(cudaMemcpyAsync(h_state, state.p, sizeof(GnState), cudaMemcpyDeviceToHost, stream));
    d2h_bytes += (long long)sizeof(GnState);
    if (log_cap) {
        FLS_CUDA(cudaMemcpyAsync(h_log.data(), log.p, sizeof(fls_iter_log) * log_cap, cudaMemcpyDeviceToHost, stream));
        d2h_bytes += (long long)(sizeof(fls_iter_log) * log_cap);
    }
    end_call(st);
    const GnState& s = *h_state;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[c * 4 + r] = s.R[r * 3 + c];
        T[12 + r] = s.t[r];
    }
    T[3] = T[7] = T[11] = 0.0;
    T[15] = 1.0;
    std::memcpy(T_final, T, sizeof(T_final));
    log_n = s.iter < log_cap ? s.iter : log_cap;
    if (converged) *converged = s.converged;
    if (st) {
        st->iterations = s.iter;
        st->converged = s.converged;
        st->n_source = n_source;
        st->n_valid = s.n_valid;
        st->sum_residual = s.sum_res;
        if (profile) {
            float tot = 0.f;
            for (int it = 0; it < s.iter && 2 * it + 1 < (int)prof_ev.size(); ++it) {
                float ms = 0.f;
                FLS_CUDA(cudaEventElapsedTime(&ms, prof_ev[2 * it], prof_ev[2 * it + 1]));
                tot += ms;
            }
            st->kernel_ms = tot;
            st->kernel_launches = s.iter;
            st->algo_bytes = (long long)s.iter * n_source * per_point_iter_bytes + (long long)(s.cand_total + 0.5) * per_cand_bytes;
        }
    }
    return FLS_OK;
}

}  // namespace fls

// =====================================================================================================================
using fls::Handle;

#define FLS_TRY try {
#define FLS_CATCH                                              \
    }                                                          \
    catch (const fls::CudaError& e) { return e.status; }       \
    catch (const std::bad_alloc&) {                            \
        fls::set_last_error("host allocation failed");         \
        return FLS_ERR_CUDA;                                   \
    }

extern "C" {

int fls_abi_version(void) { return FLS_ABI_VERSION; }

int fls_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* fls_last_error(void) { return fls::g_last_error.c_str(); }

const char* fls_strerror(int status) {
    switch (status) {
        case FLS_OK: return "ok";
        case FLS_ERR_INVALID_ARG: return "invalid argument";
        case FLS_ERR_CUDA: return "CUDA runtime error (see fls_last_error)";
        case FLS_ERR_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
        case FLS_ERR_UNSUPPORTED: return "method or mode not supported by this build";
        case FLS_ERR_NO_MAP: return "Match called before AddCloudToLocalMap";
        case FLS_ERR_CAPACITY: return "voxel capacity reached (LRU eviction is not emulated on the device)";
        case FLS_ERR_TOO_FEW_POINTS: return "too few points";
        default: return "unknown status";
    }
}

int fls_config_default(fls_config* c, int method) {
    if (!c || method < 0 || method > FLS_LOAM_FULL) return FLS_ERR_INVALID_ARG;
    std::memset(c, 0, sizeof(*c));
    c->method = method;
    c->device = 0;
    c->localization_mode = 1;
    c->max_iterations = 10;               // config/localization/config_turing.yaml:49
    c->position_converge_thres = 0.01;    // :51
    c->rotation_converge_thres = 0.01;    // :52
    c->point_to_planar_thres = 0.1;       // :50
    c->ivox_resolution = 0.5f;            // loam_point_to_plane_ivox.h:55
    c->ivox_nearby = FLS_NEARBY18;        // :56
    c->ivox_capacity = 1000000;           // ivox_map.h:35
    c->ivox_max_range = 5.0f;             // ivox_map.h:58
    c->ivox_k = 5;                        // ivox_map.h:57
    c->ndt_voxel_size = 1.0;              // config/mapping/config_nclt_ndt.yaml:42-51
    c->ndt_outlier_thres = 5.0;
    c->ndt_min_points_in_voxel = 5;
    c->ndt_max_points_in_voxel = 50;
    c->ndt_min_effective_pts = 50;
    c->ndt_capacity = 100000;
    c->icp_max_correspond_distance = 1.0;  // config/localization/config_nclt_icp.yaml:42-48
    c->rot_thre_add_cloud = 0.2;
    c->dist_thre_add_cloud = 1.0;
    c->local_map_size = 50;
    c->source_cloud_filter_size = 0.2f;
    c->map_cloud_filter_size = 0.4f;
    c->point_search_thres = 1.0;
    c->line_ratio_thres = 3.0;
    c->corner_map_filter_size = 0.2f;
    c->corner_local_map_size = 50;
    if (method == FLS_NDT) {
        c->max_iterations = 30;
        c->position_converge_thres = 0.005;
        c->rotation_converge_thres = 0.005;
    } else if (method == FLS_ICP_P2P) {
        c->max_iterations = 30;
        c->position_converge_thres = 0.005;
        c->rotation_converge_thres = 0.005;
        c->source_cloud_filter_size = 0.4f;
    } else if (method == FLS_P2PLANE_KNN) {
        c->max_iterations = 8;  // config/localization/config_nclt.yaml:45
        c->position_converge_thres = 0.005;
        c->rotation_converge_thres = 0.005;
        c->map_cloud_filter_size = 0.5f;
    } else if (method == FLS_LOAM_FULL) {
        c->max_iterations = 30;  // config/mapping/config_nclt_loam_full.yaml:40-58
        c->rotation_converge_thres = 0.05;
        c->point_to_planar_thres = 0.2;
    }
    return FLS_OK;
}

static int validate(const fls_config* c) {
    if (!c) return FLS_ERR_INVALID_ARG;
    if (c->method < 0 || c->method > FLS_LOAM_FULL) return FLS_ERR_INVALID_ARG;
    // the reference CHECK_NE()s every threshold against its "NaN" sentinel = numeric_limits::max (constant_variable.h:10-15)
    if (c->max_iterations <= 0 || c->max_iterations == 2147483647) return FLS_ERR_INVALID_ARG;
    if (!(c->position_converge_thres < 1e300) || !(c->rotation_converge_thres < 1e300)) return FLS_ERR_INVALID_ARG;
    if (c->ivox_nearby < 0 || c->ivox_nearby > 3) return FLS_ERR_INVALID_ARG;
    if (c->method == FLS_P2PLANE_IVOX) {
        if (!(c->point_to_planar_thres < 1e300) || !(c->ivox_resolution > 0.f)) return FLS_ERR_INVALID_ARG;
        if (c->ivox_k != 5) return FLS_ERR_UNSUPPORTED;  // upstream always asks for 5 (loam_point_to_plane_ivox.h:269)
    }
    return FLS_OK;
}

int fls_create(const fls_config* cfg, fls_handle** out) {
    if (!out) return FLS_ERR_INVALID_ARG;
    *out = nullptr;
    const int v = validate(cfg);
    if (v != FLS_OK) return v;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        fls::set_last_error("no usable CUDA device");
        return FLS_ERR_NO_DEVICE;
    }
    FLS_TRY
    Handle* h = new Handle(*cfg);
    *out = reinterpret_cast<fls_handle*>(h);
    return FLS_OK;
    FLS_CATCH
}

void fls_destroy(fls_handle* h) { delete reinterpret_cast<Handle*>(h); }

int fls_add_cloud(fls_handle* hh, int n_clouds, const void* const* pts, const size_t* n, size_t stride) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h || !pts || !n || n_clouds < 1 || (stride != 16 && (stride < 20 || stride % 4))) return FLS_ERR_INVALID_ARG;
    FLS_TRY
    h->begin_call();
    int rc = FLS_ERR_UNSUPPORTED;
    switch (h->cfg.method) {
        case FLS_P2PLANE_IVOX:
            if (n_clouds != 1) return FLS_ERR_INVALID_ARG;  // CHECK_EQ(cloud_list.size(), 1)
            rc = h->add_cloud_ivox(pts[0], n[0], stride);
            break;
        default: break;
    }
    h->end_call(nullptr);
    return rc;
    FLS_CATCH
}

static int match_dispatch(Handle* h, const float4* d_ordered, size_t n_ordered, const float4* d_planar, size_t n_planar, double* T,
                          int* converged, fls_match_stats* st) {
    switch (h->cfg.method) {
        case FLS_P2PLANE_IVOX: return h->match_p2plane_ivox(d_planar, n_planar, T, converged, st);
        default: return FLS_ERR_UNSUPPORTED;
    }
}

int fls_match(fls_handle* hh, const void* ordered, size_t n_ordered, const void* planar, size_t n_planar, const void* corner, size_t n_corner,
              size_t stride, double T[16], int* converged, fls_match_stats* st) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h || !T || (stride != 16 && (stride < 20 || stride % 4))) return FLS_ERR_INVALID_ARG;
    (void)corner;
    (void)n_corner;
    FLS_TRY
    if (st) std::memset(st, 0, sizeof(*st));
    h->begin_call();
    const float4* d_ord = nullptr;
    const float4* d_pla = nullptr;
    const bool uses_planar = h->cfg.method >= FLS_P2PLANE_IVOX;
    if (uses_planar) {
        if (!planar && n_planar) return FLS_ERR_INVALID_ARG;
        d_pla = h->upload(planar, n_planar, stride, h->src);
    } else {
        if (!ordered && n_ordered) return FLS_ERR_INVALID_ARG;
        d_ord = h->upload(ordered, n_ordered, stride, h->src);
    }
    return match_dispatch(h, d_ord, n_ordered, d_pla, n_planar, T, converged, st);
    FLS_CATCH
}

int fls_match_device(fls_handle* hh, const void* d_points, size_t n, double T[16], int* converged, fls_match_stats* st) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h || !T || (!d_points && n)) return FLS_ERR_INVALID_ARG;
    FLS_TRY
    if (st) std::memset(st, 0, sizeof(*st));
    h->begin_call();
    const float4* d = static_cast<const float4*>(d_points);
    return match_dispatch(h, d, n, d, n, T, converged, st);
    FLS_CATCH
}

int fls_fitness(fls_handle* hh, float max_range, float* score) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h || !score) return FLS_ERR_INVALID_ARG;
    (void)max_range;
    *score = 3.402823466e+38f;  // FloatNaN upstream (constant_variable.h:11)
    return FLS_ERR_UNSUPPORTED;
}

int fls_get_iter_log(const fls_handle* hh, fls_iter_log* out, int capacity) {
    const Handle* h = reinterpret_cast<const Handle*>(hh);
    if (!h || !out || capacity < 0) return FLS_ERR_INVALID_ARG;
    const int n = h->log_n < capacity ? h->log_n : capacity;
    for (int i = 0; i < n; ++i) out[i] = h->h_log[i];
    return n;
}

int fls_get_map_info(const fls_handle* hh, fls_map_info* out) {
    const Handle* h = reinterpret_cast<const Handle*>(hh);
    if (!h || !out) return FLS_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof(*out));
    if (h->cfg.method == FLS_P2PLANE_IVOX) {
        out->n_points = (long long)h->ivox.n_pts;
        out->n_voxels = (long long)h->ivox.n_vox;
        out->table_slots = h->ivox.n_pts ? (long long)h->ivox.mask + 1 : 0;
        out->bytes = (long long)h->ivox.bytes();
    }
    return FLS_OK;
}

int fls_ivox_knn(fls_handle* hh, const void* queries, size_t n, size_t stride, int k, float* out_pts, int32_t* out_count) {
    Handle* h = reinterpret_cast<Handle*>(hh);
    if (!h || !queries || !out_pts || !out_count || k != 5) return FLS_ERR_INVALID_ARG;
    if (h->cfg.method != FLS_P2PLANE_IVOX) return FLS_ERR_UNSUPPORTED;
    if (h->ivox.n_pts == 0) return FLS_ERR_NO_MAP;
    FLS_TRY
    h->begin_call();
    const float4* dq = h->upload(queries, n, stride, h->src);
    fls::DevBuf<float4> d_out;
    fls::DevBuf<int> d_found;
    d_out.reserve(n * 5);
    d_found.reserve(n);
    fls::launch_ivox_knn_test(h->ivox_view(), dq, (int)n, d_out.p, d_found.p, h->stream);
    FLS_CUDA(cudaMemcpyAsync(out_pts, d_out.p, n * 5 * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    FLS_CUDA(cudaMemcpyAsync(out_count, d_found.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->end_call(nullptr);
    return FLS_OK;
    FLS_CATCH
}

int fls_voxel_grid(int device, const void* pts, size_t n, size_t stride, float leaf, float* out, size_t* n_out) {
    (void)device; (void)pts; (void)n; (void)stride; (void)leaf; (void)out; (void)n_out;
    return FLS_ERR_UNSUPPORTED;
}

int fls_extract_features(const fls_feature_cfg* cfg, const float* depth, const int32_t* col, size_t n, const int32_t* row_start,
                         const int32_t* row_end, int32_t n_rows, int32_t* corner_idx, size_t* n_corner, int32_t* planar_idx, size_t* n_planar,
                         fls_match_stats* stats) {
    (void)cfg; (void)depth; (void)col; (void)n; (void)row_start; (void)row_end; (void)n_rows; (void)corner_idx; (void)n_corner; (void)planar_idx;
    (void)n_planar; (void)stats;
    return FLS_ERR_UNSUPPORTED;
}

}  // extern "C"
