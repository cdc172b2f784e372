"""include/sage_icp/pipeline/sageICP.hpp — the header the ROS 2 node recompiles against — built with g++ against minimal
Eigen/Sophus test doubles (tests/cpp/doubles; the real libraries are not in this image) and linked to the C-ABI library,
driven the way ros/ros2/OdometryServer.cpp drives the reference class."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import sage_icp_b200 as sg
    sg.build_library()
    out = str(tmp_path_factory.mktemp("adaptor") / "adaptor_check")
    libdir = os.path.join(ROOT, "sage_icp_b200", "lib")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "cpp", "doubles"), "-I",
           os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adaptor_check.cpp"), "-o", out, "-L", libdir,
           "-lsage_icp_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_adaptor_compiles_and_fails_loudly_without_a_device(exe):
    import sage_icp_b200 as sg
    if sg.device_count() > 0:
        pytest.skip("a B200 is visible here")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stdout


def test_adaptor_keeps_the_reference_surface():
    """Every public name of the reference class (pipeline/sageICP.hpp:39-99) is present with the same spelling."""
    txt = open(os.path.join(ROOT, "include", "sage_icp", "pipeline", "sageICP.hpp")).read()
    for name in ("struct sageConfig", "class sageICP", "voxel_labels", "voxel_size_map", "label_max_range", "local_map_range",
                 "basic_points_per_voxel", "critical_points_per_voxel", "basic_parts_labels", "min_motion_th", "initial_threshold",
                 "sem_th", "deskew", "dynamic_vehicle_filter_th", "dynamic_vehicle_voxid", "dynamic_remove_lankmark",
                 "Vector4dVectorTuple RegisterFrame(const std::vector<Eigen::Vector4d> &frame)",
                 "const std::vector<double> &timestamps", "Vector4dVectorTuple2 Voxelize(", "double GetAdaptiveThreshold()",
                 "Sophus::SE3d GetPredictionModel() const", "bool HasMoved()", "TransformToLastFrame(const Sophus::SE3d &last_pose",
                 "LocalMap() const", "std::vector<Sophus::SE3d> poses() const", "bool reinitialize()", "namespace sage_icp::pipeline"):
        assert name in txt, name


@pytest.mark.gpu
def test_adaptor_poses_equal_the_c_abi_path(exe, cfg, tmp_path):
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    traj = syn.trajectory(5)
    files, scans = [], []
    for i in range(5):
        scan = syn.make_scan(400 + i, tuple(traj[i]), n_beams=32, n_az=600)
        f = tmp_path / f"frame{i}.bin"
        np.ascontiguousarray(scan, np.float64).tofile(f)
        files.append(str(f)); scans.append(scan)
    r = subprocess.run([exe] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l.split() for l in r.stdout.splitlines()]
    p = sg.SagePipeline(cfg)
    for i, scan in enumerate(scans):
        pose, _, _ = p.register_frame(scan)
        got = np.array([float(v) for v in lines[i][1:8]])
        assert np.array_equal(got, pose), i  # same library, same device arithmetic: bit-identical
        assert int(lines[i][9]) == len(p.last_source())
    assert lines[5][0] == "local_map" and int(lines[5][1]) == p.map().num_points()
    assert lines[6] == ["after", "reinitialize", "poses", "0"]
