// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PCL.  Stand-ins for the PCL types the reference's Preprocess uses
// (core/Preprocessing.cpp:95-172; PCL is an unpinned system dependency there and is not in this image).
#pragma once
#include <cstdint>

namespace pcl {
struct PointXYZL {
    float x = 0, y = 0, z = 0;
    std::uint32_t label = 0;
};
}  // namespace pcl
