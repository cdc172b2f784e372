// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PCL / FLANN.
//
// pcl::KdTreeFLANN<PointT>::radiusSearch as an exhaustive search over a 0.5 m-independent uniform cell grid: every point whose
// squared float distance to the query is < radius^2, sorted by distance (index breaks ties) — what PCL returns with its default
// sorted_results = true, up to FLANN's own summation order and tie order, which cannot be known without FLANN.  Used by the
// reference only for the dynamic-vehicle filter (core/Preprocessing.cpp:126-127,150).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../point_cloud.h"
#include "../point_types.h"

namespace pcl {

template <class PointT>
class KdTreeFLANN {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr &cloud) {
        cloud_ = cloud;
        grid_.clear();
        for (std::size_t i = 0; i < cloud_->points.size(); ++i) grid_[key(cell(cloud_->points[i].x), cell(cloud_->points[i].y), cell(cloud_->points[i].z))].push_back((int)i);
    }
    int radiusSearch(const PointT &q, double radius, std::vector<int> &indices, std::vector<float> &sq_distances, unsigned max_nn = 0) const {
        indices.clear(), sq_distances.clear();
        if (!cloud_) return 0;
        const float r2 = (float)radius * (float)radius;
        const int reach = (int)std::ceil(radius / kCell) + 1;
        std::vector<std::pair<float, int>> hits;
        const long cx = cell(q.x), cy = cell(q.y), cz = cell(q.z);
        for (long x = cx - reach; x <= cx + reach; ++x)
            for (long y = cy - reach; y <= cy + reach; ++y)
                for (long z = cz - reach; z <= cz + reach; ++z) {
                    auto it = grid_.find(key(x, y, z));
                    if (it == grid_.end()) continue;
                    for (int i : it->second) {
                        const PointT &p = cloud_->points[i];
                        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                        const float d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < r2) hits.emplace_back(d2, i);
                    }
                }
        std::sort(hits.begin(), hits.end());
        if (max_nn && hits.size() > max_nn) hits.resize(max_nn);
        for (const auto &h : hits) indices.push_back(h.second), sq_distances.push_back(h.first);
        return (int)indices.size();
    }

private:
    static constexpr double kCell = 0.5;
    static long cell(float v) { return (long)std::floor((double)v / kCell); }
    static std::uint64_t key(long x, long y, long z) {
        return ((std::uint64_t)(x + (1 << 20)) & 0x1fffff) | (((std::uint64_t)(y + (1 << 20)) & 0x1fffff) << 21) | (((std::uint64_t)(z + (1 << 20)) & 0x1fffff) << 42);
    }
    typename PointCloud<PointT>::Ptr cloud_;
    std::unordered_map<std::uint64_t, std::vector<int>> grid_;
};

namespace search {
template <class PointT>
class KdTree : public ::pcl::KdTreeFLANN<PointT> {
public:
    using Ptr = std::shared_ptr<KdTree<PointT>>;
};
}  // namespace search

}  // namespace pcl
