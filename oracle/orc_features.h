// ORACLE — TEST INFRASTRUCTURE ONLY.  parity unpinned (see orc_math.h header).
//
// CPU restatement of the LOAM feature front end:
//   project()          <- src/loam/pointcloud_projector.cpp:32-133 (range image, first hit wins, row-major compaction;
//                         de-skew is the identity here: the IMU corrector is outside the hot-path scope)
//   extract_features() <- src/loam/feature_extractor.cpp:35-222
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <vector>

#include "orc_cloud.h"
#include "orc_math.h"

namespace orc {

struct Projected {
    Cloud ordered;
    std::vector<float> depth;
    std::vector<int> col;
    std::vector<int> row_start, row_end;
};

// pointcloud_projector.cpp:32-133.  raw: xyzi + ring per point, in firing order.
inline Projected project(const Cloud& raw, const std::vector<int>& ring, int V, int H, float h_res, float min_d, float max_d) {
    Projected out;
    const float FMAX = std::numeric_limits<float>::max();
    std::vector<float> range(size_t(V) * H, FMAX);
    std::vector<P4> tmp(size_t(V) * H);
    for (size_t k = 0; k < raw.size(); ++k) {
        const float x = raw[k].x, y = raw[k].y, z = raw[k].z;
        const float depth = std::sqrt(x * x + y * y + z * z);
        if (depth < min_d || depth > max_d) continue;
        const int row = ring[k];
        int colv = int(std::round(fast_atan2f(y, x) / h_res)) + H / 2;  // :68-70
        if (colv >= H) colv -= H;
        if (row >= V || row < 0 || colv < 0 || colv >= H) continue;
        const size_t index = size_t(row) * H + colv;
        if (range[index] != FMAX) continue;  // :90-91 first hit wins
        range[index] = depth;
        tmp[index] = raw[k];
    }
    out.depth.resize(size_t(V) * H);
    out.col.assign(size_t(V) * H, 0);
    out.row_start.resize(V);
    out.row_end.resize(V);
    int count = 0;
    for (int row = 0; row < V; ++row) {
        out.row_start[row] = count + 5;  // :115
        for (int c = 0; c < H; ++c) {
            const size_t index = size_t(row) * H + c;
            if (range[index] == FMAX) continue;
            out.depth[count] = range[index];
            out.ordered.push_back(tmp[index]);
            out.col[count] = c;
            ++count;
        }
        out.row_end[row] = count - 6;  // :131
    }
    return out;
}

struct Features {
    std::vector<int> corner_idx, planar_idx;  // indices into the ordered cloud, in emission order
};

// feature_extractor.cpp:35-222 on the projector's arrays.  depth/col must have at least N entries.
// std::sort(par) upstream is unstable: ties in roughness are pinned here to ascending position
// (stable sort), which the GPU path reproduces.
inline Features extract_features(int N, const float* depth, const int* col, int V, const int* row_start, const int* row_end, float corner_thr,
                                 float planar_thr) {
    Features out;
    if (N < 12) return out;
    struct PF { float rough; unsigned idx; };
    std::vector<PF> pf(size_t(N), PF{std::numeric_limits<float>::lowest(), 0u});
    std::vector<uint8_t> valid(size_t(N), 1), corner(size_t(N), 0);
    // SelectValidPoints :64-118
    for (int k = 0; k < 5; ++k) valid[k] = 0;
    for (int k = 1; k <= 6; ++k) valid[N - k] = 0;
    for (int i = 5; i < N - 6; ++i) {
        const float d1 = depth[i], d2 = depth[i + 1];
        const int cd = std::abs(col[i + 1] - col[i]);
        if (cd < 10) {
            if (d1 - d2 > 0.3) {  // float difference compared against a double literal (:89)
                for (int k = 0; k <= 5; ++k) valid[i - k] = 0;
            } else if (d2 - d1 > 0.3) {
                for (int k = 1; k <= 6; ++k) valid[i + k] = 0;
            }
        }
        const float diff1 = std::abs(depth[i - 1] - depth[i]);
        const float diff2 = std::abs(depth[i + 1] - depth[i]);
        if (diff1 > 0.02 * depth[i] && diff2 > 0.02 * depth[i]) valid[i] = 0;  // double product (:113)
    }
    // ComputeRoughness :46-61
    for (int i = 5; i < N - 5; ++i) {
        const float r = depth[i - 5] + depth[i - 4] + depth[i - 3] + depth[i - 2] + depth[i - 1] + depth[i + 1] + depth[i + 2] + depth[i + 3] +
                        depth[i + 4] + depth[i + 5] - 10.0f * depth[i];
        pf[i].rough = r * r;
        pf[i].idx = unsigned(i);
    }
    // SelectFeatures :120-222
    auto suppress = [&](int index) {
        for (int k = 1; k <= 5; ++k) {
            if (std::abs(col[index + k] - col[index + k - 1]) > 10) break;
            valid[index + k] = 0;
        }
        for (int k = -1; k >= -5; --k) {
            if (std::abs(col[index + k] - col[index + k + 1]) > 10) break;
            valid[index + k] = 0;
        }
    };
    for (int scan = 0; scan < V; ++scan) {
        for (int b = 0; b < 6; ++b) {
            const int len = (row_end[scan] - row_start[scan]) / 6;
            const int bs = row_start[scan] + b * len, be = row_start[scan] + (b + 1) * len;
            if (bs >= be) continue;
            std::stable_sort(pf.begin() + bs, pf.begin() + be, [](const PF& l, const PF& r) { return l.rough < r.rough; });
            int picked = 0;
            for (int j = be; j >= bs; --j) {  // inclusive of `be`  [quirk 10]
                const int index = int(pf[j].idx);
                if (pf[j].rough > corner_thr && valid[index]) {
                    picked++;
                    if (picked <= 20) {
                        corner[index] = 1;
                        out.corner_idx.push_back(index);
                    } else {
                        break;
                    }
                    valid[index] = 0;
                    suppress(index);
                }
            }
            for (int j = bs; j <= be; ++j) {  // inclusive of `be`  [quirk 10]
                const int index = int(pf[j].idx);
                if (valid[index] && pf[j].rough < planar_thr) {
                    valid[index] = 0;
                    suppress(index);
                }
                if (!corner[index]) out.planar_idx.push_back(index);  // every non-corner, regardless of the threshold (:214-216)
            }
        }
    }
    return out;
}

}  // namespace orc
