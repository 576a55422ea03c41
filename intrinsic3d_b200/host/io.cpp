// File formats and pose conversions around the refinement path: intrinsics text file (src/camera.cpp:202-274),
// TUM trajectory files (src/rgbd/sensor.cpp:236-347), pose vector <-> matrix (src/math.cpp:151-178).  Pure host code.
#include <nv/camera.h>
#include <nv/math.h>

#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace nv
{
bool Camera::load(const std::string& filename)
{
    std::ifstream in(filename.c_str());
    bool loaded = false;
    if (in.is_open())
    {
        int w = 0, h = 0;
        float K[9], d[5];
        bool ok = static_cast<bool>(in >> w >> h);
        for (int i = 0; i < 9 && ok; ++i) ok = static_cast<bool>(in >> K[i]);
        for (int i = 0; i < 5 && ok; ++i) ok = static_cast<bool>(in >> d[i]);
        if (ok)
        {
            width_ = w; height_ = h;
            fx_ = K[0]; fy_ = K[4]; cx_ = K[2]; cy_ = K[5];
            for (int i = 0; i < 5; ++i) dist_[i] = d[i];
            loaded = true;
        }
    }
    if (!loaded)
    {
        std::cout << "Intrinsics file ('" << filename << "') could not be loaded! Using defaults..." << std::endl;
        const int w = width_, h = height_;
        setDefault();
        if (w > 0 && h > 0) { width_ = w; height_ = h; }
    }
    return loaded;
}

bool Camera::save(const std::string& filename) const
{
    if (filename.empty()) return false;
    std::ofstream out(filename.c_str());
    if (!out.is_open()) return false;
    out << width_ << " " << height_ << std::endl;
    out << static_cast<float>(fx_) << " 0 " << static_cast<float>(cx_) << std::endl;
    out << "0 " << static_cast<float>(fy_) << " " << static_cast<float>(cy_) << std::endl;
    out << "0 0 1" << std::endl;
    out << static_cast<float>(dist_[0]) << " " << static_cast<float>(dist_[1]) << " " << static_cast<float>(dist_[2]) << " " << static_cast<float>(dist_[3]) << " "
        << static_cast<float>(dist_[4]) << std::endl;
    return out.good();
}

namespace math
{
Mat4 poseVecAAToMat(const Vec6& p)
{
    Mat4 M;
    const double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const double angle = std::sqrt(n2);
    double ax = p[0], ay = p[1], az = p[2];
    if (n2 > 0.0) { ax /= angle; ay /= angle; az /= angle; }        // Eigen normalized(): unchanged when the norm is 0
    // Eigen::AngleAxis::toRotationMatrix
    const double s = std::sin(angle), c = std::cos(angle);
    const double sx = s * ax, sy = s * ay, sz = s * az;
    const double c1x = (1.0 - c) * ax, c1y = (1.0 - c) * ay, c1z = (1.0 - c) * az;
    double t;
    t = c1x * ay; M(0, 1) = t - sz; M(1, 0) = t + sz;
    t = c1x * az; M(0, 2) = t + sy; M(2, 0) = t - sy;
    t = c1y * az; M(1, 2) = t - sx; M(2, 1) = t + sx;
    M(0, 0) = c1x * ax + c; M(1, 1) = c1y * ay + c; M(2, 2) = c1z * az + c;
    M(0, 3) = p[3]; M(1, 3) = p[4]; M(2, 3) = p[5];
    return M;
}

Vec6 poseMatToVecAA(const Mat4& M)
{
    // Eigen::AngleAxisd(rotation matrix) goes through a quaternion
    double qw, qx, qy, qz;
    const double tr = M(0, 0) + M(1, 1) + M(2, 2);
    if (tr > 0.0)
    {
        double t = std::sqrt(tr + 1.0);
        qw = 0.5 * t; t = 0.5 / t;
        qx = (M(2, 1) - M(1, 2)) * t; qy = (M(0, 2) - M(2, 0)) * t; qz = (M(1, 0) - M(0, 1)) * t;
    }
    else
    {
        int i = 0;
        if (M(1, 1) > M(0, 0)) i = 1;
        if (M(2, 2) > M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
        double q[3];
        q[i] = 0.5 * t; t = 0.5 / t;
        qw = (M(k, j) - M(j, k)) * t;
        q[j] = (M(j, i) + M(i, j)) * t; q[k] = (M(k, i) + M(i, k)) * t;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    // quaternion -> angle axis (Eigen::AngleAxis::operator=(QuaternionBase)): angle in [0, pi]
    double n = std::sqrt(qx * qx + qy * qy + qz * qz);
    Vec6 v;
    if (n != 0.0)
    {
        const double angle = 2.0 * std::atan2(n, std::fabs(qw));
        if (qw < 0.0) n = -n;
        v[0] = qx / n * angle; v[1] = qy / n * angle; v[2] = qz / n * angle;
    }
    else { v[0] = v[1] = v[2] = 0.0; }
    v[3] = M(0, 3); v[4] = M(1, 3); v[5] = M(2, 3);
    return v;
}

Mat4 invertPose(const Mat4& P)
{
    Mat4 I;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) I(r, c) = P(c, r);
    for (int r = 0; r < 3; ++r) I(r, 3) = -(I(r, 0) * P(0, 3) + I(r, 1) * P(1, 3) + I(r, 2) * P(2, 3));
    return I;
}
} // namespace math

bool loadPoses(const std::string& filename, std::vector<Mat4f>& poses, std::vector<double>& timestamps, bool first_pose_is_identity)
{
    // like the reference (src/rgbd/sensor.cpp:235-279) the poses are APPENDED to what the vectors already hold; the optional
    // re-basing below applies to the appended range only when the vectors were empty (the only way the reference calls it)
    std::ifstream in(filename.c_str());
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line))
    {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream iss(line);
        double ts; float tx, ty, tz, qx, qy, qz, qw;
        if (!(iss >> ts >> tx >> ty >> tz >> qx >> qy >> qz >> qw)) break;
        timestamps.push_back(ts);
        Mat4f p;
        // Eigen::Quaternionf(qw, qx, qy, qz).toRotationMatrix() (no normalisation)
        const float tx2 = 2.0f * qx, ty2 = 2.0f * qy, tz2 = 2.0f * qz;
        const float twx = tx2 * qw, twy = ty2 * qw, twz = tz2 * qw;
        const float txx = tx2 * qx, txy = ty2 * qx, txz = tz2 * qx;
        const float tyy = ty2 * qy, tyz = tz2 * qy, tzz = tz2 * qz;
        p(0, 0) = 1.0f - (tyy + tzz); p(0, 1) = txy - twz; p(0, 2) = txz + twy;
        p(1, 0) = txy + twz; p(1, 1) = 1.0f - (txx + tzz); p(1, 2) = tyz - twx;
        p(2, 0) = txz - twy; p(2, 1) = tyz + twx; p(2, 2) = 1.0f - (txx + tyy);
        p(0, 3) = tx; p(1, 3) = ty; p(2, 3) = tz;
        poses.push_back(p);
    }
    if (first_pose_is_identity && !poses.empty())
    {
        // poses[i] = poses[0]^-1 * poses[i]
        const Mat4f P0 = poses[0];
        for (Mat4f& P : poses)
        {
            Mat4f R;
            for (int r = 0; r < 3; ++r)
            {
                for (int c = 0; c < 3; ++c) R(r, c) = P0(0, r) * P(0, c) + P0(1, r) * P(1, c) + P0(2, r) * P(2, c);
                R(r, 3) = P0(0, r) * (P(0, 3) - P0(0, 3)) + P0(1, r) * (P(1, 3) - P0(1, 3)) + P0(2, r) * (P(2, 3) - P0(2, 3));
            }
            P = R;
        }
    }
    return true;
}

bool savePoses(const std::string& filename, const std::vector<Mat4f>& poses, const std::vector<double>& timestamps)
{
    if (filename.empty() || poses.size() != timestamps.size()) return false;
    std::ofstream out(filename.c_str());
    if (!out.is_open()) return false;
    out << std::fixed << std::setprecision(6);
    for (size_t i = 0; i < poses.size(); ++i)
    {
        const Mat4f& M = poses[i];
        // Eigen::Quaternionf(rotation matrix)
        float qw, qx, qy, qz;
        const float tr = M(0, 0) + M(1, 1) + M(2, 2);
        if (tr > 0.0f)
        {
            float t = std::sqrt(tr + 1.0f);
            qw = 0.5f * t; t = 0.5f / t;
            qx = (M(2, 1) - M(1, 2)) * t; qy = (M(0, 2) - M(2, 0)) * t; qz = (M(1, 0) - M(0, 1)) * t;
        }
        else
        {
            int a = 0;
            if (M(1, 1) > M(0, 0)) a = 1;
            if (M(2, 2) > M(a, a)) a = 2;
            const int b = (a + 1) % 3, c = (b + 1) % 3;
            float t = std::sqrt(M(a, a) - M(b, b) - M(c, c) + 1.0f);
            float q[3];
            q[a] = 0.5f * t; t = 0.5f / t;
            qw = (M(c, b) - M(b, c)) * t;
            q[b] = (M(b, a) + M(a, b)) * t; q[c] = (M(c, a) + M(a, c)) * t;
            qx = q[0]; qy = q[1]; qz = q[2];
        }
        out << timestamps[i] << " " << M(0, 3) << " " << M(1, 3) << " " << M(2, 3) << " " << qx << " " << qy << " " << qz << " " << qw << std::endl;
    }
    return out.good();
}
} // namespace nv
