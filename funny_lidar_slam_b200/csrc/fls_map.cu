// fls_map.cu — device-side construction of the point grids (iVox map, ICP / fitness search grids): K8, batch form.
//
// IVoxMap::AddPoints (src/ivox_map/ivox_map.cpp:122-143 upstream) inserts points one by one into an
// unordered_map of std::list nodes.  Here a whole cloud is inserted at once:
//   key (Morton of round(p/res)) -> stable radix sort -> gather -> run-length encode -> scan -> hash insert,
// then (iVox only) the per-centre stencil lists are materialised (see fls_ivox.cuh).
// The stable sort keeps insertion order inside a voxel, so the k-NN tie order matches a sequential insert.
// LRU eviction (capacity_) is not emulated: a build that would reach the capacity returns FLS_ERR_CAPACITY
// (DESIGN.md "Scope"); upstream evicts nothing while size() < capacity_.
#include <cub/cub.cuh>

#include "fls_ivox.cuh"
#include "fls_maps.h"

namespace fls {

BuildScratch::BuildScratch() { cudaMallocHost(&h_num_runs, sizeof(int)); }
BuildScratch::~BuildScratch() {
    if (h_num_runs) cudaFreeHost(h_num_runs);
}

namespace {

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }

__global__ void ivox_keys_kernel(const float4* __restrict__ pts, size_t n, float inv_res, int key_mode, unsigned long long* __restrict__ keys,
                                 unsigned* __restrict__ idx) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    keys[i] = morton_key(grid_coord(p.x, inv_res, key_mode), grid_coord(p.y, inv_res, key_mode), grid_coord(p.z, inv_res, key_mode));
    idx[i] = (unsigned)i;
}

__global__ void gather_kernel(const float4* __restrict__ src, const unsigned* __restrict__ idx, size_t n, float4* __restrict__ dst) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void table_clear_kernel(HashSlot* tab, size_t slots) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < slots) {
        tab[i].key = kEmptyKey;
        tab[i].start = 0;
        tab[i].count = 0;
    }
}

__device__ __forceinline__ void table_insert(HashSlot* tab, unsigned mask, unsigned long long key, unsigned start, unsigned count) {
    unsigned h = hash_key(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&tab[h].key, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            tab[h].start = start;
            tab[h].count = count;
            return;
        }
        h = (h + 1) & mask;
    }
}

// one thread per occupied voxel (run of equal Morton codes)
__global__ void ivox_insert_kernel(const unsigned long long* __restrict__ run_morton, const unsigned* __restrict__ starts,
                                   const unsigned* __restrict__ counts, int n_runs, HashSlot* tab, unsigned mask) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_runs) return;
    int x, y, z;
    morton_decode(run_morton[v], x, y, z);
    table_insert(tab, mask, pack_key(x, y, z), starts[v], counts[v]);
}

// ---- stencil lists ---------------------------------------------------------------------------------------------------
__global__ void center_keys_kernel(const unsigned long long* __restrict__ run_morton, int n_runs, int n_stencil,
                                   unsigned long long* __restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)n_runs * n_stencil) return;
    const int v = (int)(t / n_stencil), s = (int)(t % n_stencil);
    int x, y, z;
    morton_decode(run_morton[v], x, y, z);
    // the stencil is symmetric: voxel v is in the stencil of centre c  <=>  c = v - offset
    out[t] = morton_key(x - c_stencil[s][0], y - c_stencil[s][1], z - c_stencil[s][2]);
}

__global__ void list_count_kernel(const unsigned long long* __restrict__ centers, int n_centers, int n_stencil, const HashSlot* __restrict__ tab,
                                  unsigned mask, unsigned* __restrict__ ccount) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_centers) return;
    int x, y, z;
    morton_decode(centers[c], x, y, z);
    unsigned tot = 0;
    for (int s = 0; s < n_stencil; ++s) {
        unsigned st, cnt;
        if (table_find(tab, mask, pack_key(x + c_stencil[s][0], y + c_stencil[s][1], z + c_stencil[s][2]), st, cnt)) tot += cnt;
    }
    ccount[c] = tot;
}

// one warp per centre: concatenate the points of its stencil voxels in visit order; lane 0 publishes the centre slot
__global__ void list_fill_kernel(const unsigned long long* __restrict__ centers, int n_centers, int n_stencil, const HashSlot* __restrict__ tab,
                                 unsigned mask, const float4* __restrict__ pts, const unsigned* __restrict__ cstart,
                                 const unsigned* __restrict__ ccount, float4* __restrict__ lists, HashSlot* ctab, unsigned cmask) {
    const int c = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n_centers) return;
    int x, y, z;
    morton_decode(centers[c], x, y, z);
    unsigned off = cstart[c];
    for (int s = 0; s < n_stencil; ++s) {
        unsigned st, cnt;
        if (!table_find(tab, mask, pack_key(x + c_stencil[s][0], y + c_stencil[s][1], z + c_stencil[s][2]), st, cnt)) continue;
        for (unsigned k = lane; k < cnt; k += 32) lists[off + k] = pts[st + k];
        off += cnt;
    }
    if (lane == 0) table_insert(ctab, cmask, pack_key(x, y, z), cstart[c], ccount[c]);
}

__global__ void repack_kernel(const unsigned char* __restrict__ raw, size_t n, size_t stride, float4* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* f = reinterpret_cast<const float*>(raw + i * stride);
    float4 p;
    p.x = f[0];
    p.y = f[1];
    p.z = f[2];
    p.w = f[4];  // pcl::PointXYZI keeps intensity at byte 16
    out[i] = p;
}

__global__ void transform_f_kernel(const float4* __restrict__ in, size_t n, float r0, float r1, float r2, float r3, float r4, float r5, float r6,
                                   float r7, float r8, float t0, float t1, float t2, float4* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    out[i] = make_float4(xform_row_f(r0, r1, r2, t0, p.x, p.y, p.z), xform_row_f(r3, r4, r5, t1, p.x, p.y, p.z),
                         xform_row_f(r6, r7, r8, t2, p.x, p.y, p.z), p.w);
}

// pcl::transformPoint / pcl::transformPointCloud with a double transform: fp64 R·p + t, stored back as fp32
__global__ void transform_d_kernel(const float4* __restrict__ in, size_t n, double r0, double r1, double r2, double r3, double r4, double r5,
                                   double r6, double r7, double r8, double t0, double t1, double t2, float4* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    out[i] = make_float4(xform_row_d(r0, r1, r2, t0, (double)p.x, (double)p.y, (double)p.z),
                         xform_row_d(r3, r4, r5, t1, (double)p.x, (double)p.y, (double)p.z),
                         xform_row_d(r6, r7, r8, t2, (double)p.x, (double)p.y, (double)p.z), p.w);
}

}  // namespace

void launch_transform_d(const float4* d_in, size_t n, const double* T, float4* d_out, cudaStream_t st) {
    if (n == 0) return;
    transform_d_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_in, n, T[0], T[4], T[8], T[1], T[5], T[9], T[2], T[6], T[10], T[12], T[13], T[14], d_out);
}

void launch_repack(const unsigned char* d_raw, size_t n, size_t stride, float4* d_out, cudaStream_t st) {
    if (n == 0) return;
    repack_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_raw, n, stride, d_out);
}

void launch_transform_f(const float4* d_in, size_t n, const double* T, float4* d_out, cudaStream_t st) {
    if (n == 0) return;
    // T column-major: R(r,c) = T[c*4+r]
    transform_f_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_in, n, (float)T[0], (float)T[4], (float)T[8], (float)T[1], (float)T[5], (float)T[9],
                                                        (float)T[2], (float)T[6], (float)T[10], (float)T[12], (float)T[13], (float)T[14], d_out);
}

int IvoxMap::append_and_build(const float4* d_new, size_t n_new, long long capacity, cudaStream_t st) {
    const size_t n_old = n_pts;
    const size_t n = n_old + n_new;
    if (n == 0) return FLS_OK;
    if (n > 0xfffffff0ull) return FLS_ERR_INVALID_ARG;
    // grow pts_all preserving the old contents
    if (n > pts_all.cap) {
        DevBuf<float4> bigger;
        bigger.reserve(n + n / 2);
        if (n_old) FLS_CUDA(cudaMemcpyAsync(bigger.p, pts_all.p, n_old * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        std::swap(bigger.p, pts_all.p);
        std::swap(bigger.cap, pts_all.cap);
    }
    if (n_new) FLS_CUDA(cudaMemcpyAsync(pts_all.p + n_old, d_new, n_new * sizeof(float4), cudaMemcpyDeviceToDevice, st));

    BuildScratch& sc = scratch;
    sc.keys.reserve(n);
    sc.keys_sorted.reserve(n);
    sc.uniq.reserve(n);
    sc.idx.reserve(n);
    sc.idx_sorted.reserve(n);
    sc.counts.reserve(n);
    sc.starts.reserve(n);
    sc.num_runs.reserve(2);
    pts_sorted.reserve(n);

    ivox_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(pts_all.p, n, inv_res, key_mode, sc.keys.p, sc.idx.p);
    size_t tmp1 = 0, tmp2 = 0, tmp3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp1, sc.keys.p, sc.keys_sorted.p, sc.idx.p, sc.idx_sorted.p, (int)n, 0, 63, st);
    cub::DeviceRunLengthEncode::Encode(nullptr, tmp2, sc.keys_sorted.p, sc.uniq.p, sc.counts.p, sc.num_runs.p, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp3, sc.counts.p, sc.starts.p, (int)n, st);
    size_t tmp = tmp1 > tmp2 ? tmp1 : tmp2;
    tmp = tmp > tmp3 ? tmp : tmp3;
    sc.cub_tmp.reserve(tmp + 256);
    size_t tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortPairs(sc.cub_tmp.p, tb, sc.keys.p, sc.keys_sorted.p, sc.idx.p, sc.idx_sorted.p, (int)n, 0, 63, st));
    gather_kernel<<<grid_for(n, 256), 256, 0, st>>>(pts_all.p, sc.idx_sorted.p, n, pts_sorted.p);
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRunLengthEncode::Encode(sc.cub_tmp.p, tb, sc.keys_sorted.p, sc.uniq.p, sc.counts.p, sc.num_runs.p, (int)n, st));
    FLS_CUDA(cudaMemcpyAsync(sc.h_num_runs, sc.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const int runs = *sc.h_num_runs;
    if (capacity > 0 && (long long)runs >= capacity) return FLS_ERR_CAPACITY;
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceScan::ExclusiveSum(sc.cub_tmp.p, tb, sc.counts.p, sc.starts.p, runs, st));
    size_t slots = 1024;
    while (slots < 2 * (size_t)runs) slots <<= 1;
    table.reserve(slots);
    mask = (unsigned)(slots - 1);
    table_clear_kernel<<<grid_for(slots, 256), 256, 0, st>>>(table.p, slots);
    ivox_insert_kernel<<<grid_for(runs, 256), 256, 0, st>>>(sc.uniq.p, sc.starts.p, sc.counts.p, runs, table.p, mask);
    FLS_CUDA(cudaGetLastError());
    n_pts = n;
    n_vox = (size_t)runs;
    launches += 8;
    if (n_stencil > 0) return build_stencil_lists(st);
    return FLS_OK;
}

int IvoxMap::build_stencil_lists(cudaStream_t st) {
    BuildScratch& sc = scratch;
    const size_t S = (size_t)n_stencil;
    const size_t n_keys = n_vox * S;
    const size_t total = n_pts * S;  // every point lands in exactly S lists (symmetric stencil)
    if (total > 0xfffffff0ull || n_keys > 0x7ffffff0ull) return FLS_ERR_CAPACITY;
    ckeys.reserve(n_keys);
    ckeys_sorted.reserve(n_keys);
    cuniq.reserve(n_keys);
    center_keys_kernel<<<grid_for(n_keys, 256), 256, 0, st>>>(sc.uniq.p, (int)n_vox, n_stencil, ckeys.p);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, t1, ckeys.p, ckeys_sorted.p, (int)n_keys, 0, 63, st);
    cub::DeviceSelect::Unique(nullptr, t2, ckeys_sorted.p, cuniq.p, sc.num_runs.p, (int)n_keys, st);
    sc.cub_tmp.reserve((t1 > t2 ? t1 : t2) + 256);
    size_t tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortKeys(sc.cub_tmp.p, tb, ckeys.p, ckeys_sorted.p, (int)n_keys, 0, 63, st));
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceSelect::Unique(sc.cub_tmp.p, tb, ckeys_sorted.p, cuniq.p, sc.num_runs.p, (int)n_keys, st));
    FLS_CUDA(cudaMemcpyAsync(sc.h_num_runs, sc.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const int nc = *sc.h_num_runs;
    ccount.reserve((size_t)nc);
    cstart.reserve((size_t)nc);
    list_count_kernel<<<grid_for((size_t)nc, 128), 128, 0, st>>>(cuniq.p, nc, n_stencil, table.p, mask, ccount.p);
    size_t t3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t3, ccount.p, cstart.p, nc, st);
    sc.cub_tmp.reserve(t3 + 256);
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceScan::ExclusiveSum(sc.cub_tmp.p, tb, ccount.p, cstart.p, nc, st));
    // load factor <= 0.25: the centre table is probed once per point-iteration and a long linear-probing chain stalls
    // a whole warp, so it is kept sparser than the occupied table
    size_t slots = 1024;
    while (slots < 4 * (size_t)nc) slots <<= 1;
    ctab.reserve(slots);
    cmask = (unsigned)(slots - 1);
    lists.reserve(total);
    table_clear_kernel<<<grid_for(slots, 256), 256, 0, st>>>(ctab.p, slots);
    list_fill_kernel<<<grid_for((size_t)nc * 32, 256), 256, 0, st>>>(cuniq.p, nc, n_stencil, table.p, mask, pts_sorted.p, cstart.p, ccount.p, lists.p,
                                                                     ctab.p, cmask);
    FLS_CUDA(cudaGetLastError());
    n_centers = (size_t)nc;
    n_list = total;
    launches += 7;
    // the transient key arrays are the largest buffers of the build; give them back
    ckeys.release();
    ckeys_sorted.release();
    return FLS_OK;
}

}  // namespace fls
