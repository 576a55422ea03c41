// nv/mat.h — minimal fixed-size vector PODs carrying the names the reference gets from Eigen
// (libintrinsic3d/include/nv/mat.h:51-85).  Eigen is not available in this build environment; a maintainer integrating
// into the real tree keeps the Eigen typedefs — the shims only use operator[], data() and size().
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace nv
{
template <class T, int N>
struct VecN
{
    std::array<T, N> v{};
    VecN() = default;
    VecN(std::initializer_list<T> l) { int i = 0; for (T x : l) { if (i < N) v[i++] = x; } }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    static constexpr int size() { return N; }
    bool operator==(const VecN& o) const { return v == o.v; }
    VecN operator+(const VecN& o) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    VecN operator*(T s) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * s; return r; }
    static VecN Zero() { return VecN(); }
};
using Vec3i = VecN<int, 3>;
using Vec3f = VecN<float, 3>;
using Vec3b = VecN<unsigned char, 3>;
using Vec4 = VecN<double, 4>;
using Vec5 = VecN<double, 5>;
using Vec6 = VecN<double, 6>;
using VecXd = std::vector<double>;   // stands in for Eigen::VectorXd (per-voxel SH coefficients)
// 4x4 double matrix (Eigen::Matrix4d): poses in math::poseVecAAToMat / poseMatToVecAA
struct Mat4
{
    double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double& operator()(int r, int c) { return m[4 * r + c]; }
    double operator()(int r, int c) const { return m[4 * r + c]; }
    static Mat4 Identity() { return Mat4(); }
};
// 4x4 float matrix (Eigen::Matrix4f in the reference): the world->camera pose SDFColorization::add receives
struct Mat4f
{
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float& operator()(int r, int c) { return m[4 * r + c]; }
    float operator()(int r, int c) const { return m[4 * r + c]; }
    static Mat4f Identity() { return Mat4f(); }
};
} // namespace nv

namespace std
{
// spatial hash of integer voxel coordinates (Teschner et al.), the role of mat.h:115-124 in the reference
template <>
struct hash<nv::Vec3i>
{
    size_t operator()(const nv::Vec3i& p) const noexcept
    {
        return (static_cast<size_t>(static_cast<uint32_t>(p[0])) * 73856093u) ^ (static_cast<size_t>(static_cast<uint32_t>(p[1])) * 19349669u) ^
               (static_cast<size_t>(static_cast<uint32_t>(p[2])) * 83492791u);
    }
};
} // namespace std
