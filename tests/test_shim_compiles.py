"""The header-only adapter (funny_lidar_slam_b200/shim/b200_registration.h) only compiles inside the reference's build.
Here it is compiled against minimal stand-ins of the reference headers it includes — same names, same member names, same
virtual signatures as include/registration/registration_interface.h:11-20 and include/lidar/pointcloud_cluster.h:13-26
upstream — so that a signature drift between the adapter, the C ABI and the interface is caught on CPU."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MOCKS = {
    "glog/logging.h": """
#pragma once
#include <iostream>
struct NullStream { template <class T> NullStream& operator<<(const T&) { return *this; } };
#define CHECK_EQ(a, b) ((a) == (b) ? NullStream() : NullStream())
#define LOG(x) NullStream()
#define DLOG(x) NullStream()
""",
    "common/constant_variable.h": """
#pragma once
#include <limits>
static const float FloatNaN = std::numeric_limits<float>::max();
""",
    "common/data_type.h": """
#pragma once
#include <memory>
#include <vector>
struct alignas(16) PCLPointXYZI { float x, y, z, pad; float intensity, p1, p2, p3; };
struct PCLPointCloudXYZI { std::vector<PCLPointXYZI> points; size_t size() const { return points.size(); } };
struct Mat4d { double m[16]; double* data() { return m; } const double* data() const { return m; } };
struct PointcloudCluster { PCLPointCloudXYZI ordered_cloud_, planar_cloud_, corner_cloud_; };
using PointcloudClusterPtr = std::shared_ptr<PointcloudCluster>;
""",
    "registration/registration_interface.h": """
#pragma once
#include <initializer_list>
#include "common/data_type.h"
class RegistrationInterface {
public:
    virtual bool Match(const PointcloudClusterPtr& source_cloud_cluster, Mat4d& T) = 0;
    virtual void AddCloudToLocalMap(const std::initializer_list<PCLPointCloudXYZI>& cloud_list) = 0;
    [[nodiscard]] virtual float GetFitnessScore(float max_range) const = 0;
    virtual ~RegistrationInterface() = default;
};
""",
}

USER = """
#include <string>
#include "b200_registration.h"
int use(const fls_config& cfg) {
    std::shared_ptr<RegistrationInterface> m = B200Registration::Create("PointToPlane_IVOX", cfg);
    auto cluster = std::make_shared<PointcloudCluster>();
    Mat4d T{};
    PCLPointCloudXYZI planar, corner;
    m->AddCloudToLocalMap({planar});
    m->AddCloudToLocalMap({planar, corner});
    const bool ok = m->Match(cluster, T);
    return ok ? (m->GetFitnessScore(2.0f) < 1.0f) : 0;
}
"""


def test_adapter_compiles_against_the_reference_interface(tmp_path):
    gxx = shutil.which("g++") or "/usr/bin/g++"
    if not os.path.exists(gxx):
        return
    for rel, body in MOCKS.items():
        p = tmp_path / "mock" / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(body)
    (tmp_path / "user.cpp").write_text(USER)
    cmd = [gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fsyntax-only", "-I", str(tmp_path / "mock"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "funny_lidar_slam_b200", "shim"), str(tmp_path / "user.cpp")]
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
