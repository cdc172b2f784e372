// Compile/link check of include/sage_icp/pipeline/sageICP.hpp exactly the way ros/ros2/OdometryServer.cpp uses it
// (member by value, assigned from sageICP(config): OdometryServer.hpp:86, OdometryServer.cpp:104,167,173,218,263).
// Reads frames as raw doubles from stdin-style files given on the command line, prints poses; with no argument it only
// exercises construction (on a box without a B200 that must throw, loudly).
#include <cstdio>
#include <fstream>
#include <iostream>

#include "sage_icp/pipeline/sageICP.hpp"

using sage_icp::pipeline::sageConfig;
using sage_icp::pipeline::sageICP;

static sageConfig launch_config() {  // ros/launch/odometry.launch.py:32-67, dynamic filter off
    sageConfig c;
    c.voxel_labels = {{40, 44, 48, 49}, {50, 51, 52}, {70, 72}, {60, 71, 80, 81, 99}, {0}, {10, 11, 13, 15, 16, 18, 20}};
    c.voxel_size = {0.6, 1.0, 0.9, 0.8, 1.0, 0.6};
    c.voxel_size_map = 0.8;
    c.basic_parts_labels = {40, 44, 48, 49, 50, 70, 72};
    return c;
}

struct Node {  // the shape of sage_icp_ros::OdometryServer
    sageICP odometry_;
    sageConfig config_;
};

int main(int argc, char **argv) {
    Node node;
    node.config_ = launch_config();
    try {
        node.odometry_ = sageICP(node.config_);
    } catch (const std::exception &e) {
        std::printf("construction failed: %s\n", e.what());
        return 3;
    }
    for (int a = 1; a < argc; ++a) {
        std::ifstream f(argv[a], std::ios::binary | std::ios::ate);
        const size_t n = static_cast<size_t>(f.tellg()) / (4 * sizeof(double));
        f.seekg(0);
        std::vector<Eigen::Vector4d> frame(n);
        f.read(reinterpret_cast<char *>(frame.data()), static_cast<std::streamsize>(n * 4 * sizeof(double)));
        const std::vector<double> timestamps;  // deskew is off in every launch file
        const auto &[source, t_icp, t_all] = node.odometry_.RegisterFrame(frame, timestamps);
        const auto pose = node.odometry_.poses().back();
        const auto t = pose.translation();
        const auto q = pose.unit_quaternion();
        std::printf("pose %.17g %.17g %.17g %.17g %.17g %.17g %.17g source %zu t_icp %.6f t_all %.6f\n", t[0], t[1], t[2], q.x(), q.y(),
                    q.z(), q.w(), source.size(), t_icp, t_all);
    }
    if (argc > 1) {
        const auto map = node.odometry_.LocalMap();
        const auto [src, ds] = node.odometry_.Voxelize(map);
        std::printf("local_map %zu voxelized %zu %zu has_moved %d\n", map.size(), src.size(), ds.size(), (int)node.odometry_.HasMoved());
        node.odometry_.reinitialize();
        std::printf("after reinitialize poses %zu\n", node.odometry_.poses().size());
    }
    return 0;
}
