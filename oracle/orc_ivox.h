// ORACLE — TEST INFRASTRUCTURE ONLY.  parity unpinned (see orc_math.h header).
//
// CPU restatement of the incremental voxel map: include/ivox_map/ivox_map.h:16-74,
// src/ivox_map/ivox_map.cpp:6-37 (GetClosestPoint), :43-66 (stencils), :122-143 (AddPoints + LRU),
// :145-147 (Pos2Grid), src/ivox_map/voxel_grid_node.cpp:23-42 (per-voxel bounded k-NN).
#pragma once
#include <list>
#include <unordered_map>
#include <vector>

#include "orc_cloud.h"

namespace orc {

enum NearbyType { NEARBY_CENTER = 0, NEARBY6 = 1, NEARBY18 = 2, NEARBY26 = 3 };

// stencil offsets in the reference's order (ivox_map.cpp:43-66)
inline const int (*ivox_stencil(int type, int* n))[3] {
    static const int S[27][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 1}, {1, 1, 0}, {-1, 1, 0},
                                 {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1},
                                 {0, -1, -1}, {1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};
    static const int counts[4] = {1, 7, 19, 27};
    *n = counts[type];
    return S;
}

class IVox {
public:
    IVox(float resolution, int nearby_type, size_t capacity) : res_(resolution), inv_res_(1.0f / resolution), type_(nearby_type), capacity_(capacity) {}

    // ivox_map.cpp:145-147 — key = round(p * inv_res) per axis, fp32 product, std::round (half away from zero)
    void pos2grid(const P4& p, int* k) const {
        k[0] = int(std::round(p.x * inv_res_));
        k[1] = int(std::round(p.y * inv_res_));
        k[2] = int(std::round(p.z * inv_res_));
    }

    // ivox_map.cpp:122-143 — sequential insert, move-to-front on touch, evict the LRU tail when
    // size() >= capacity after creating a voxel.
    void add_points(const Cloud& pts) {
        for (const P4& p : pts) {
            int k[3];
            pos2grid(p, k);
            const uint64_t key = pack(k);
            auto it = map_.find(key);
            if (it == map_.end()) {
                cache_.push_front(Node{key, {}});
                map_[key] = cache_.begin();
                cache_.front().pts.push_back(p);
                if (map_.size() >= capacity_) {
                    map_.erase(cache_.back().key);
                    cache_.pop_back();
                }
            } else {
                it->second->pts.push_back(p);
                cache_.splice(cache_.begin(), cache_, it->second);
                map_[key] = cache_.begin();
            }
        }
    }

    // ivox_map.cpp:6-37 + voxel_grid_node.cpp:23-42.  Candidates = points with d2 < max_range^2 in the
    // stencil voxels; each voxel contributes at most K (its K nearest); global K nearest kept; the
    // nearest is moved to slot 0.  std::nth_element leaves tie / residual order unspecified upstream —
    // the oracle pins it: ascending (d2, visit order).
    int closest(const P4& q, int K, float max_range, P4* out) const {
        struct Cand { double d; const P4* p; int ord; };  // same 24-byte record as upstream's DistPoint
        auto less = [](const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.ord < b.ord); };
        int ns;
        const int(*S)[3] = ivox_stencil(type_, &ns);
        std::vector<Cand> cands;
        cands.reserve(size_t(K) * ns);
        int k[3];
        pos2grid(q, k);
        int ord = 0;
        for (int s = 0; s < ns; ++s) {
            const int kk[3] = {k[0] + S[s][0], k[1] + S[s][1], k[2] + S[s][2]};
            auto it = map_.find(pack(kk));
            if (it == map_.end()) continue;
            const size_t old = cands.size();
            for (const P4& p : it->second->pts) {
                const double d = double(dist2f(p, q));
                if (d < double(max_range * max_range)) cands.push_back({d, &p, ord++});
            }
            if (old + K < cands.size()) {  // voxel_grid_node.cpp:33-39
                std::nth_element(cands.begin() + old, cands.begin() + old + K - 1, cands.end(), less);
                cands.resize(old + K);
            }
        }
        if (cands.empty()) return 0;
        if (int(cands.size()) > K) {  // ivox_map.cpp:25-28
            std::nth_element(cands.begin(), cands.begin() + K - 1, cands.end(), less);
            cands.resize(K);
        }
        std::sort(cands.begin(), cands.end(), less);  // upstream only moves the minimum to slot 0 (:30); full order pinned here
        for (size_t i = 0; i < cands.size(); ++i) out[i] = *cands[i].p;
        return int(cands.size());
    }

    size_t num_voxels() const { return map_.size(); }
    size_t num_points() const {
        size_t n = 0;
        for (const auto& nd : cache_) n += nd.pts.size();
        return n;
    }
    // flat dump (voxel order = LRU order, front first) for tests
    void dump(Cloud& out) const {
        for (const auto& nd : cache_) out.insert(out.end(), nd.pts.begin(), nd.pts.end());
    }

private:
    struct Node { uint64_t key; std::vector<P4> pts; };
    static uint64_t pack(const int* k) {
        return (uint64_t(uint32_t(k[0]) & 0x1fffff) << 42) | (uint64_t(uint32_t(k[1]) & 0x1fffff) << 21) | uint64_t(uint32_t(k[2]) & 0x1fffff);
    }
    float res_, inv_res_;
    int type_;
    size_t capacity_;
    std::list<Node> cache_;
    std::unordered_map<uint64_t, std::list<Node>::iterator> map_;
};

}  // namespace orc
