// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PCL (see point_types.h).
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

namespace pcl {
template <class PointT>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    std::size_t size() const { return points.size(); }
};
struct PointIndices {
    std::vector<int> indices;
};
}  // namespace pcl
