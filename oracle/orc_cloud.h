// ORACLE — TEST INFRASTRUCTURE ONLY.  parity unpinned (see orc_math.h header).
//
// Point-cloud primitives on the hot path that live in un-vendored third-party code upstream
// (PCL 1.10 as shipped by osrf/ros:noetic, Dockerfile:1): pcl::VoxelGrid::filter,
// pcl::KdTreeFLANN::nearestKSearch (exact k-NN), pcl::transformPoint; plus the repo's own
// float transform (include/common/pointcloud_utility.h:52-72,141-158).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace orc {

struct P4 {  // x, y, z, intensity — the 16 payload bytes of pcl::PointXYZI
    float x, y, z, i;
};
using Cloud = std::vector<P4>;

// fp32 squared distance, evaluation order (dx*dx + dy*dy) + dz*dz, no FMA contraction
// (include/common/pointcloud_utility.h:13-17; FLANN L2_Simple<float> accumulates in the same order).
inline float dist2f(const P4& a, const P4& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

// pcl::transformPoint(pt, Eigen::Transform<double,3,Affine>) — loam_point_to_plane_ivox.h:91,265:
// coordinates promoted to double, ((r0*x + r1*y) + r2*z) + t, rounded once to float.
// T is 4x4 row-major double.
inline P4 transform_point_d(const P4& p, const double* T) {
    P4 o;
    o.x = float(((T[0] * double(p.x) + T[1] * double(p.y)) + T[2] * double(p.z)) + T[3]);
    o.y = float(((T[4] * double(p.x) + T[5] * double(p.y)) + T[6] * double(p.z)) + T[7]);
    o.z = float(((T[8] * double(p.x) + T[9] * double(p.y)) + T[10] * double(p.z)) + T[11]);
    o.i = p.i;
    return o;
}

// TransformPoint(pt, Mat3f R, Vec3f t) with R,t cast to float FIRST (pointcloud_utility.h:63-72,145-146).
inline P4 transform_point_f(const P4& p, const float* Rf /*3x3 row-major*/, const float* tf) {
    P4 o;
    o.x = ((Rf[0] * p.x + Rf[1] * p.y) + Rf[2] * p.z) + tf[0];
    o.y = ((Rf[3] * p.x + Rf[4] * p.y) + Rf[5] * p.z) + tf[1];
    o.z = ((Rf[6] * p.x + Rf[7] * p.y) + Rf[8] * p.z) + tf[2];
    o.i = p.i;
    return o;
}
inline Cloud transform_cloud_f(const Cloud& c, const double* T) {
    float Rf[9], tf[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rf[i * 3 + j] = float(T[i * 4 + j]);
        tf[i] = float(T[i * 4 + 3]);
    }
    Cloud o(c.size());
    for (size_t k = 0; k < c.size(); ++k) o[k] = transform_point_f(c[k], Rf, tf);
    return o;
}

// pcl::VoxelGrid<PointXYZI>::filter with downsample_all_data_=true, min_points_per_voxel_=0, as wrapped
// by VoxelGridCloud (pointcloud_utility.h:216-224,263-271).  PCL 1.10 semantics restated:
//   * bounding box over all points; inverse leaf = 1/leaf in fp32;
//   * if dx*dy*dz > INT_MAX the input is returned unchanged;
//   * cell = floor(x*inv) - min_b per axis (fp32), linear id = i + j*dx + k*dx*dy;
//   * points sorted by id — PCL uses std::sort (unstable); the oracle pins ties to input order
//     (stable), which fixes the fp32 summation order of each centroid;
//   * one output per occupied cell, fp32 running sum of xyz and intensity divided by float(n),
//     cells emitted in ascending id.
inline Cloud voxel_grid(const Cloud& in, float leaf) {
    Cloud out;
    if (in.empty()) return out;
    const float inv = 1.0f / leaf;
    float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
    for (const P4& p : in) {
        mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
        mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
        mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = int64_t((mx[0] - mn[0]) * inv) + 1, dy = int64_t((mx[1] - mn[1]) * inv) + 1, dz = int64_t((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > int64_t(INT_MAX)) return in;
    int minb[3], maxb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
        minb[a] = int(std::floor(mn[a] * inv));
        maxb[a] = int(std::floor(mx[a] * inv));
        divb[a] = maxb[a] - minb[a] + 1;
    }
    const int mul1 = divb[0], mul2 = divb[0] * divb[1];
    std::vector<std::pair<unsigned, unsigned>> iv(in.size());
    for (size_t k = 0; k < in.size(); ++k) {
        const int i0 = int(std::floor(in[k].x * inv) - float(minb[0]));
        const int i1 = int(std::floor(in[k].y * inv) - float(minb[1]));
        const int i2 = int(std::floor(in[k].z * inv) - float(minb[2]));
        iv[k] = {unsigned(i0 + i1 * mul1 + i2 * mul2), unsigned(k)};
    }
    std::stable_sort(iv.begin(), iv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    size_t s = 0;
    while (s < iv.size()) {
        size_t e = s + 1;
        while (e < iv.size() && iv[e].first == iv[s].first) ++e;
        float sx = 0, sy = 0, sz = 0, si = 0;
        for (size_t k = s; k < e; ++k) {
            const P4& p = in[iv[k].second];
            sx += p.x; sy += p.y; sz += p.z; si += p.i;
        }
        const float n = float(e - s);
        out.push_back({sx / n, sy / n, sz / n, si / n});
        s = e;
    }
    return out;
}

// Exact k-nearest-neighbour index standing in for pcl::KdTreeFLANN<PointXYZI>::nearestKSearch
// (icp_optimized.h:85,203; loam_point_to_plane_kdtree.h:217; loam_full_kdtree.h:225,289):
// exact search, fp32 squared L2, ascending, ties to the lower index.  Implemented as a uniform
// grid with an expanding shell search (the result is exact, checked against scipy.cKDTree in tests).
class ExactKnn {
public:
    void build(const Cloud& pts, float cell = 1.0f) {
        pts_ = pts;
        cell_ = cell;
        inv_ = 1.0f / cell;
        grid_.clear();
        for (size_t k = 0; k < pts_.size(); ++k) grid_[key(ci(pts_[k].x), ci(pts_[k].y), ci(pts_[k].z))].push_back(int(k));
        if (!pts_.empty()) {
            lo_[0] = hi_[0] = ci(pts_[0].x); lo_[1] = hi_[1] = ci(pts_[0].y); lo_[2] = hi_[2] = ci(pts_[0].z);
            for (const P4& p : pts_) {
                const int c[3] = {ci(p.x), ci(p.y), ci(p.z)};
                for (int a = 0; a < 3; ++a) { lo_[a] = std::min(lo_[a], c[a]); hi_[a] = std::max(hi_[a], c[a]); }
            }
        }
    }
    size_t size() const { return pts_.size(); }
    const P4& point(int i) const { return pts_[i]; }

    // returns number found (min(k, size)); idx/d2 sorted ascending
    int search(const P4& q, int k, int* idx, float* d2) const {
        const int n = int(std::min<size_t>(k, pts_.size()));
        if (n == 0) return 0;
        std::vector<std::pair<float, int>> best;  // kept sorted, size <= n
        const int c[3] = {ci(q.x), ci(q.y), ci(q.z)};
        int maxshell = 0;
        for (int a = 0; a < 3; ++a) maxshell = std::max(maxshell, std::max(std::abs(c[a] - lo_[a]), std::abs(hi_[a] - c[a])));
        bool done = false;
        for (int r = 0; r <= maxshell; ++r) {
            if (int(best.size()) == n) {
                // every point in shell r is at least (r-1)*cell away along some axis
                const float lb = float(r - 1) * cell_;
                if (r >= 1 && lb > 0 && lb * lb > best.back().first) { done = true; break; }
            }
            if (r > 6) break;  // sparse / far query: shells grow as r^3 — finish by exhaustive scan below
            for (int dx = -r; dx <= r; ++dx)
                for (int dy = -r; dy <= r; ++dy)
                    for (int dz = -r; dz <= r; ++dz) {
                        if (std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz))) != r) continue;
                        auto it = grid_.find(key(c[0] + dx, c[1] + dy, c[2] + dz));
                        if (it == grid_.end()) continue;
                        for (int pi : it->second) {
                            const float d = dist2f(pts_[pi], q);
                            std::pair<float, int> e{d, pi};
                            if (int(best.size()) < n) {
                                best.insert(std::upper_bound(best.begin(), best.end(), e), e);
                            } else if (e < best.back()) {
                                best.pop_back();
                                best.insert(std::upper_bound(best.begin(), best.end(), e), e);
                            }
                        }
                    }
        }
        if (!done && maxshell > 6) {  // exhaustive, exact by construction
            best.clear();
            for (size_t pi = 0; pi < pts_.size(); ++pi) {
                std::pair<float, int> e{dist2f(pts_[pi], q), int(pi)};
                if (int(best.size()) < n) {
                    best.insert(std::upper_bound(best.begin(), best.end(), e), e);
                } else if (e < best.back()) {
                    best.pop_back();
                    best.insert(std::upper_bound(best.begin(), best.end(), e), e);
                }
            }
        }
        for (int i = 0; i < n; ++i) { idx[i] = best[i].second; d2[i] = best[i].first; }
        return n;
    }

private:
    int ci(float v) const { return int(std::floor(v * inv_)); }
    static uint64_t key(int x, int y, int z) {
        return (uint64_t(uint32_t(x) & 0x1fffff) << 42) | (uint64_t(uint32_t(y) & 0x1fffff) << 21) | uint64_t(uint32_t(z) & 0x1fffff);
    }
    Cloud pts_;
    float cell_ = 1.0f, inv_ = 1.0f;
    int lo_[3] = {0, 0, 0}, hi_[3] = {0, 0, 0};
    std::unordered_map<uint64_t, std::vector<int>> grid_;
};

}  // namespace orc
