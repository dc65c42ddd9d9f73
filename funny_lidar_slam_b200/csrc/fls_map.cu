// fls_map.cu — device-side construction of the point grids (iVox map, ICP / fitness search grids): K8, batch form.
//
// IVoxMap::AddPoints (src/ivox_map/ivox_map.cpp:122-143 upstream) inserts points one by one into an
// unordered_map of std::list nodes.  Here a whole cloud is inserted at once:
//   key (Morton of round(p/res)) -> stable radix sort -> gather -> run-length encode -> scan -> hash insert,
// then (iVox only) the per-centre stencil lists are materialised (see fls_ivox.cuh).
// The stable sort keeps insertion order inside a voxel, so the k-NN tie order matches a sequential insert.
// LRU eviction (capacity_, ivox_map.cpp:133-136) is emulated exactly: every point carries its insertion stamp, the stamp of a
// voxel's last point is its position in upstream's list, and the sequential insert of a call is simulated on the host against
// the candidates (IvoxMap::evict_lru, lru_simulate).
#include <cub/cub.cuh>

#include <cstdlib>
#include <functional>
#include <queue>
#include <vector>

#include "fls_ivox.cuh"
#include "fls_maps.h"

namespace fls {

bool lru_simulate(size_t size0, size_t capacity, const std::vector<unsigned>& cand_first_touch, std::vector<unsigned> create_times,
                  std::vector<unsigned>& victims, std::vector<unsigned char>& recreated) {
    victims.clear();
    recreated.clear();
    // creation events in time order; a victim that is touched later in the call adds one (it is created again, empty)
    std::priority_queue<unsigned, std::vector<unsigned>, std::greater<unsigned>> events(std::greater<unsigned>(), std::move(create_times));
    size_t size = size0, ci = 0;
    while (!events.empty()) {
        const unsigned t = events.top();
        events.pop();
        ++size;
        if (size < capacity) continue;
        // pop_back(): the oldest voxel that has not been moved to the front by an earlier point of this call
        while (ci < cand_first_touch.size() && cand_first_touch[ci] < t) ++ci;
        if (ci >= cand_first_touch.size()) return false;
        const unsigned ft = cand_first_touch[ci];
        victims.push_back((unsigned)ci);
        recreated.push_back(ft != 0xffffffffu ? 1 : 0);
        if (ft != 0xffffffffu) events.push(ft);
        ++ci;
        --size;
    }
    return true;
}

BuildScratch::BuildScratch() {
    if (cudaMallocHost(&h_num_runs, sizeof(int)) != cudaSuccess) {  // no device / out of pinned memory: a plain allocation still works as a copy target
        cudaGetLastError();
        h_num_runs = static_cast<int*>(std::malloc(sizeof(int)));
        pinned = false;
    }
}
BuildScratch::~BuildScratch() {
    if (h_num_runs) {
        if (pinned) cudaFreeHost(h_num_runs);
        else std::free(h_num_runs);
    }
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

// ---- LRU bookkeeping (ivox_map.cpp:122-143) ----------------------------------------------------------------------------
__global__ void ivox_stamp_kernel(unsigned long long* __restrict__ stamps, size_t n, unsigned long long call_hi) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) stamps[i] = call_hi | (unsigned long long)i;
}
// per voxel of the sorted order (old points first inside a voxel: the sort is stable and the new points have the highest
// indices): how many points it held before the call, the stamp of the last of them (its LRU position), the first point of this
// call that touches it.  cnt[0] candidates (voxels that existed), cnt[1] creations, cnt[2] existing voxels touched by the call.
__global__ void ivox_run_info_kernel(int runs, const unsigned* __restrict__ starts, const unsigned* __restrict__ counts,
                                     const unsigned* __restrict__ idx_sorted, const unsigned long long* __restrict__ stamp_all, unsigned n_old,
                                     unsigned* __restrict__ nold_out, unsigned* __restrict__ first_out, unsigned long long* __restrict__ cand_stamp,
                                     unsigned* __restrict__ cand_run, unsigned* __restrict__ create_times, int* cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= runs) return;
    const unsigned s = starts[r], c = counts[r];
    unsigned lo = 0, hi = c;  // first position whose point is new
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (idx_sorted[s + mid] < n_old) lo = mid + 1;
        else hi = mid;
    }
    const unsigned first = lo < c ? idx_sorted[s + lo] - n_old : 0xffffffffu;
    nold_out[r] = lo;
    first_out[r] = first;
    if (lo > 0) {
        const int pos = atomicAdd(cnt, 1);
        cand_stamp[pos] = stamp_all[idx_sorted[s + lo - 1]];
        cand_run[pos] = (unsigned)r;
        if (lo < c) atomicAdd(cnt + 2, 1);
    } else {
        create_times[atomicAdd(cnt + 1, 1)] = first;
    }
}
__global__ void ivox_cand_kernel(const unsigned* __restrict__ runs_sorted, int K, const unsigned* __restrict__ first, unsigned* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) out[k] = first[runs_sorted[k]];
}
// one warp per victim: its points from before the call are dropped
__global__ void ivox_kill_kernel(const unsigned* __restrict__ victim_pos, int n_victims, const unsigned* __restrict__ runs_sorted,
                                 const unsigned* __restrict__ starts, const unsigned* __restrict__ nold, const unsigned* __restrict__ idx_sorted,
                                 unsigned char* __restrict__ keep) {
    const int v = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= n_victims) return;
    const unsigned r = runs_sorted[victim_pos[v]];
    const unsigned s = starts[r], m = nold[r];
    for (unsigned k = lane; k < m; k += 32) keep[idx_sorted[s + k]] = 0;
}
__global__ void table_dump_kernel(const HashSlot* __restrict__ tab, size_t slots, unsigned long long* __restrict__ out, int* cursor) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < slots && tab[i].key != kEmptyKey) out[atomicAdd(cursor, 1)] = tab[i].key;
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

// keys -> stable sort -> gather -> run-length encode -> starts: the voxel-contiguous order of the first n points of pts_all
int IvoxMap::sort_and_runs(size_t n, cudaStream_t st, int* runs_out) {
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
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceScan::ExclusiveSum(sc.cub_tmp.p, tb, sc.counts.p, sc.starts.p, runs, st));
    launches += 6;
    *runs_out = runs;
    return FLS_OK;
}

int IvoxMap::append_and_build(const float4* d_new, size_t n_new, long long capacity, cudaStream_t st) {
    const size_t n_old = n_pts;
    size_t n = n_old + n_new;
    if (n == 0) return FLS_OK;
    if (n > 0xfffffff0ull) return FLS_ERR_INVALID_ARG;
    const bool lru = capacity > 0;  // the iVox map proper (the search grids have no capacity)
    // grow pts_all (and the insertion stamps) preserving the old contents
    if (n > pts_all.cap) {
        DevBuf<float4> bigger;
        bigger.reserve(n + n / 2);
        if (n_old) FLS_CUDA(cudaMemcpyAsync(bigger.p, pts_all.p, n_old * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        std::swap(bigger.p, pts_all.p);
        std::swap(bigger.cap, pts_all.cap);
    }
    if (lru && n > stamp_all.cap) {
        DevBuf<unsigned long long> bigger;
        bigger.reserve(n + n / 2);
        if (n_old) FLS_CUDA(cudaMemcpyAsync(bigger.p, stamp_all.p, n_old * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        FLS_CUDA(cudaStreamSynchronize(st));
        std::swap(bigger.p, stamp_all.p);
        std::swap(bigger.cap, stamp_all.cap);
    }
    if (n_new) FLS_CUDA(cudaMemcpyAsync(pts_all.p + n_old, d_new, n_new * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    if (lru && n_new) {
        ++call_no;
        ivox_stamp_kernel<<<grid_for(n_new, 256), 256, 0, st>>>(stamp_all.p + n_old, n_new, call_no << 32);
    }
    int runs = 0;
    int rc = sort_and_runs(n, st, &runs);
    if (rc != FLS_OK) return rc;
    if (lru && (long long)runs >= capacity) {
        // IVoxMap::AddPoints would have evicted the LRU tail while inserting (ivox_map.cpp:133-136): drop those voxels' old points
        size_t n_after = n;
        rc = evict_lru(n_old, n, runs, capacity, st, &n_after);
        if (rc != FLS_OK) return rc;
        n = n_after;
        rc = sort_and_runs(n, st, &runs);
        if (rc != FLS_OK) return rc;
    }
    size_t slots = 1024;
    while (slots < 2 * (size_t)runs) slots <<= 1;
    table.reserve(slots);
    mask = (unsigned)(slots - 1);
    table_clear_kernel<<<grid_for(slots, 256), 256, 0, st>>>(table.p, slots);
    ivox_insert_kernel<<<grid_for(runs, 256), 256, 0, st>>>(scratch.uniq.p, scratch.starts.p, scratch.counts.p, runs, table.p, mask);
    FLS_CUDA(cudaGetLastError());
    n_pts = n;
    n_vox = (size_t)runs;
    launches += 2;
    if (n_stencil > 0) return build_stencil_lists(st);
    return FLS_OK;
}

// Exact LRU of IVoxMap::AddPoints for this call (see lru_simulate): a voxel's position in upstream's list is the insertion time of
// its last point, so the stamps of the points are all the state there is.  Victims lose every point they held before the call;
// one that is touched again later in the call keeps this call's points (it is created anew).  Compacts pts_all / stamp_all.
int IvoxMap::evict_lru(size_t n_old, size_t n, int runs, long long capacity, cudaStream_t st, size_t* n_after) {
    BuildScratch& sc = scratch;
    if (n_vox == 0) return FLS_ERR_CAPACITY;  // the first cloud alone overflows the capacity
    lru_old.reserve((size_t)runs + 1);
    lru_first.reserve((size_t)runs + 1);
    lru_nold.reserve((size_t)runs + 1);
    lru_keys.reserve((size_t)runs + 1);
    lru_keys_sorted.reserve((size_t)runs + 1);
    lru_vals.reserve((size_t)runs + 1);
    lru_vals_sorted.reserve((size_t)runs + 1);
    lru_cnt.reserve(4);
    FLS_CUDA(cudaMemsetAsync(lru_cnt.p, 0, 4 * sizeof(int), st));
    ivox_run_info_kernel<<<grid_for(runs, 256), 256, 0, st>>>(runs, sc.starts.p, sc.counts.p, sc.idx_sorted.p, stamp_all.p, (unsigned)n_old, lru_nold.p,
                                                            lru_first.p, lru_keys.p, lru_vals.p, sc.k32b.reserve(n + 1), lru_cnt.p);
    int hc[4] = {0, 0, 0, 0};
    FLS_CUDA(cudaMemcpyAsync(hc, lru_cnt.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const int n_cand = hc[0], n_create = hc[1], n_touched = hc[2];
    if (n_cand == 0) return FLS_ERR_CAPACITY;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, lru_keys.p, lru_keys_sorted.p, lru_vals.p, lru_vals_sorted.p, n_cand, 0, 64, st);
    sc.cub_tmp.reserve(tb + 256);
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceRadixSort::SortPairs(sc.cub_tmp.p, tb, lru_keys.p, lru_keys_sorted.p, lru_vals.p, lru_vals_sorted.p, n_cand, 0, 64, st));
    const long long e0 = (long long)n_vox + n_create - (capacity - 1);
    size_t K = (size_t)(e0 > 0 ? e0 : 0) + 2 * (size_t)n_touched + 64;
    if (K > (size_t)n_cand) K = (size_t)n_cand;
    sc.k32a.reserve(K + 1);
    ivox_cand_kernel<<<grid_for(K, 256), 256, 0, st>>>(lru_vals_sorted.p, (int)K, lru_first.p, sc.k32a.p);
    std::vector<unsigned> cand(K), creat((size_t)n_create);
    FLS_CUDA(cudaMemcpyAsync(cand.data(), sc.k32a.p, sizeof(unsigned) * K, cudaMemcpyDeviceToHost, st));
    if (n_create) FLS_CUDA(cudaMemcpyAsync(creat.data(), sc.k32b.p, sizeof(unsigned) * (size_t)n_create, cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned> victims;
    std::vector<unsigned char> recreated;
    if (!lru_simulate(n_vox, (size_t)capacity, cand, creat, victims, recreated)) return FLS_ERR_CAPACITY;
    launches += 4;
    *n_after = n;
    if (victims.empty()) return FLS_OK;
    // flags: 1 = keep; the victims' points from before this call go
    lru_flags.reserve(n + 1);
    FLS_CUDA(cudaMemsetAsync(lru_flags.p, 1, n, st));
    sc.k32a.reserve(victims.size() + 1);
    FLS_CUDA(cudaMemcpyAsync(sc.k32a.p, victims.data(), sizeof(unsigned) * victims.size(), cudaMemcpyHostToDevice, st));
    ivox_kill_kernel<<<grid_for(victims.size() * 32, 256), 256, 0, st>>>(sc.k32a.p, (int)victims.size(), lru_vals_sorted.p, sc.starts.p, lru_nold.p,
                                                                        sc.idx_sorted.p, lru_flags.p);
    // stable compaction of the points and their stamps (pts_sorted / keys are rebuilt by the second sort anyway: use them as targets)
    sc.keys.reserve(n + 1);
    size_t t1 = 0, t2 = 0;
    cub::DeviceSelect::Flagged(nullptr, t1, pts_all.p, lru_flags.p, pts_sorted.p, sc.num_runs.p, (int)n, st);
    cub::DeviceSelect::Flagged(nullptr, t2, stamp_all.p, lru_flags.p, sc.keys.p, sc.num_runs.p, (int)n, st);
    sc.cub_tmp.reserve((t1 > t2 ? t1 : t2) + 256);
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceSelect::Flagged(sc.cub_tmp.p, tb, pts_all.p, lru_flags.p, pts_sorted.p, sc.num_runs.p, (int)n, st));
    tb = sc.cub_tmp.cap;
    FLS_CUDA(cub::DeviceSelect::Flagged(sc.cub_tmp.p, tb, stamp_all.p, lru_flags.p, sc.keys.p, sc.num_runs.p, (int)n, st));
    FLS_CUDA(cudaMemcpyAsync(sc.h_num_runs, sc.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));  // also: the host vectors above are read by the copies
    const size_t kept = (size_t)*sc.h_num_runs;
    FLS_CUDA(cudaMemcpyAsync(pts_all.p, pts_sorted.p, kept * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    FLS_CUDA(cudaMemcpyAsync(stamp_all.p, sc.keys.p, kept * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    launches += 4;
    *n_after = kept;
    return FLS_OK;
}

size_t IvoxMap::dump_keys(unsigned long long* h_out, size_t cap, cudaStream_t st) {
    // packed keys of the occupied voxels, from the table (tests)
    if (n_vox == 0) return 0;
    const size_t slots = (size_t)mask + 1;
    scratch.keys.reserve(n_vox + 1);
    scratch.num_runs.reserve(2);
    FLS_CUDA(cudaMemsetAsync(scratch.num_runs.p, 0, sizeof(int), st));
    table_dump_kernel<<<grid_for(slots, 256), 256, 0, st>>>(table.p, slots, scratch.keys.p, scratch.num_runs.p);
    int n = 0;
    FLS_CUDA(cudaMemcpyAsync(&n, scratch.num_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLS_CUDA(cudaStreamSynchronize(st));
    const size_t m = (size_t)n < cap ? (size_t)n : cap;
    FLS_CUDA(cudaMemcpy(h_out, scratch.keys.p, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost));
    return m;
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
