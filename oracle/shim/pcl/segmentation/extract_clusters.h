// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PCL.
//
// pcl::EuclideanClusterExtraction restated from PCL's published algorithm (pcl/segmentation/impl/extract_clusters.hpp):
// region growing over radius searches from every unprocessed point, a cluster is kept if min <= size <= max, its indices are
// sorted, and the clusters are finally ordered by descending size.  Equal-sized clusters keep discovery order here (PCL uses an
// unstable std::sort, so their order there is unspecified).  Used by the reference only for the dynamic-vehicle filter
// (core/Preprocessing.cpp:129-140).
#pragma once
#include <algorithm>
#include <vector>

#include "../kdtree/kdtree_flann.h"

namespace pcl {

template <class PointT>
class EuclideanClusterExtraction {
public:
    void setClusterTolerance(double t) { tolerance_ = t; }
    void setMinClusterSize(int n) { min_ = n; }
    void setMaxClusterSize(int n) { max_ = n; }
    void setSearchMethod(const typename search::KdTree<PointT>::Ptr &tree) { tree_ = tree; }
    void setInputCloud(const typename PointCloud<PointT>::Ptr &cloud) { cloud_ = cloud; }
    void extract(std::vector<PointIndices> &clusters) {
        clusters.clear();
        if (!cloud_ || !tree_) return;
        const std::size_t n = cloud_->points.size();
        std::vector<char> processed(n, 0);
        std::vector<int> nn;
        std::vector<float> nd;
        for (std::size_t i = 0; i < n; ++i) {
            if (processed[i]) continue;
            std::vector<int> queue{(int)i};
            processed[i] = 1;
            for (std::size_t s = 0; s < queue.size(); ++s) {
                if (!tree_->radiusSearch(cloud_->points[queue[s]], tolerance_, nn, nd)) continue;
                for (int j : nn)
                    if (!processed[j]) processed[j] = 1, queue.push_back(j);
            }
            if ((int)queue.size() >= min_ && (int)queue.size() <= max_) {
                PointIndices r;
                r.indices = queue;
                std::sort(r.indices.begin(), r.indices.end());
                clusters.push_back(std::move(r));
            }
        }
        std::stable_sort(clusters.begin(), clusters.end(), [](const PointIndices &a, const PointIndices &b) { return a.indices.size() > b.indices.size(); });
    }

private:
    double tolerance_ = 0;
    int min_ = 1, max_ = 1 << 30;
    typename search::KdTree<PointT>::Ptr tree_;
    typename PointCloud<PointT>::Ptr cloud_;
};

}  // namespace pcl
