/*
 * oracle.cpp — CPU oracle for the Intrinsic3D joint-refinement hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (intrinsic3d_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference (NVlabs/intrinsic3d @ 05312e5) ships no tests,
 * fixtures or golden vectors, and cannot be built here (needs Ceres 2.1.0,
 * Eigen, OpenCV, Boost — none present, no network).  This file is a float64
 * restatement of the reference algorithm and of the Ceres 2.1.0 pieces it
 * calls, written from the behaviour of the sources cited below; it is pinned
 * only by the known-answer tests in tests/ (finite differences, an independent
 * torch-fp64 autograd restatement, dense normal-equation solves).
 *
 * What is restated (paths relative to /root/reference/libintrinsic3d):
 *   Optimizer::optimize/addVoxelResiduals/buildProblem/fixVoxelParams
 *                                      src/refinement/optimizer.cpp:109-361
 *   NLSSolver::addResidual/buildProblem/normalizeCostTermWeights/solve
 *                                      src/refinement/nls_solver.cpp:172-394
 *   ShadingCost functor + create       include/nv/refinement/shading_cost.h:85-198
 *                                      src/refinement/shading_cost.cpp:59-150
 *   helpers                            include/nv/refinement/cost.h:73-127
 *                                      include/nv/sdf/operators.h:49-109
 *                                      src/sdf/operators.cpp:45-77,142-147
 *                                      include/nv/shading.h:53-148
 *                                      include/nv/camera.h:92-126 (CameraT)
 *   regularisers                       volumetric_regularizer.{h,cpp}, surface_stab_regularizer.{h,cpp},
 *                                      albedo_regularizer.{h,cpp}
 *   observation selection              src/sdf/colorization.cpp:192-370, src/camera.cpp:124-154,
 *                                      src/math.cpp:43-47,151-163
 *   grid predicates                    src/sparse_voxel_grid.cpp:166-259, src/sdf/algorithms.cpp:75-91,240-247
 *   grid-level transitions             src/sdf/algorithms.cpp:118-235 (interpolate, upsample), :368-458 (clearVoxelsOutsideThinShell)
 *   voxel recolouring                  src/sdf/colorization.cpp:113-189,215-251,318-370, src/rgbd/processing.cpp:236-302,
 *                                      src/refinement/intrinsic3d.cpp:381-409 (Intrinsic3D::recomputeColors)
 *   SVSH lighting                      src/lighting/lighting_svsh.cpp:93-346, src/lighting/subvolumes.cpp:66-304,
 *                                      src/math.cpp:74-128 (average, interpolationWeights), include/nv/shading.h:53-91
 * Ceres 2.1.0 semantics restated from memory of the upstream sources (not
 * available offline): ScaledLoss, constant parameter blocks, forward-mode Jets,
 * AngleAxisRotatePoint, BiCubicInterpolator/Grid2D, TrustRegionMinimizer with
 * LevenbergMarquardtStrategy, CgnrSolver + BlockJacobiPreconditioner +
 * ConjugateGradientsSolver (Q-based termination), Jacobi column scaling.
 *
 * Canonical order: the reference iterates a std::unordered_map; here "iteration
 * order" is the order of the flat input arrays (voxel index).  Top-K frame
 * selection breaks weight ties by frame id (higher id kept), the definition the
 * GPU engine matches bit-for-bit.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <array>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/i3d_types.h"

namespace
{

// ---------------------------------------------------------------------------
// forward-mode dual numbers, 4 derivative lanes (DynamicAutoDiffCostFunction<.,4>,
// src/refinement/shading_cost.cpp:85)
// ---------------------------------------------------------------------------
constexpr int kStride = 4;

struct Jet
{
    double a;
    double v[kStride];
    Jet() : a(0.0) { for (int i = 0; i < kStride; ++i) v[i] = 0.0; }
    Jet(double s) : a(s) { for (int i = 0; i < kStride; ++i) v[i] = 0.0; }
};

inline Jet operator+(const Jet& x, const Jet& y) { Jet r; r.a = x.a + y.a; for (int i = 0; i < kStride; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
inline Jet operator-(const Jet& x, const Jet& y) { Jet r; r.a = x.a - y.a; for (int i = 0; i < kStride; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
inline Jet operator-(const Jet& x) { Jet r; r.a = -x.a; for (int i = 0; i < kStride; ++i) r.v[i] = -x.v[i]; return r; }
inline Jet operator*(const Jet& x, const Jet& y) { Jet r; r.a = x.a * y.a; for (int i = 0; i < kStride; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
inline Jet operator/(const Jet& x, const Jet& y)
{
    Jet r; const double inv = 1.0 / y.a; const double q = x.a * inv; r.a = q;
    for (int i = 0; i < kStride; ++i) r.v[i] = (x.v[i] - q * y.v[i]) * inv;
    return r;
}
inline Jet& operator+=(Jet& x, const Jet& y) { x = x + y; return x; }
inline Jet jsqrt(const Jet& x) { Jet r; r.a = std::sqrt(x.a); const double t = 1.0 / (2.0 * r.a); for (int i = 0; i < kStride; ++i) r.v[i] = x.v[i] * t; return r; }
inline Jet jsin(const Jet& x) { Jet r; r.a = std::sin(x.a); const double c = std::cos(x.a); for (int i = 0; i < kStride; ++i) r.v[i] = c * x.v[i]; return r; }
inline Jet jcos(const Jet& x) { Jet r; r.a = std::cos(x.a); const double s = -std::sin(x.a); for (int i = 0; i < kStride; ++i) r.v[i] = s * x.v[i]; return r; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jcos(double x) { return std::cos(x); }
inline double scalar(const Jet& x) { return x.a; }
inline double scalar(double x) { return x; }
inline bool finite_all(double x) { return std::isfinite(x); }
inline bool finite_all(const Jet& x)
{
    // ceres::IsNaN / IsInfinite on a Jet look at the value and every derivative lane
    if (!std::isfinite(x.a)) return false;
    for (int i = 0; i < kStride; ++i) if (!std::isfinite(x.v[i])) return false;
    return true;
}

// ---------------------------------------------------------------------------
// image sampling: ceres::BiCubicInterpolator over Grid2D<float,1,true,true>
// (include/nv/refinement/cost.h:108-127).  Returns f, df/drow, df/dcol.
// ---------------------------------------------------------------------------
inline void cubic_hermite(double p0, double p1, double p2, double p3, double x, double* f, double* dfdx)
{
    const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
    const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
    const double c = 0.5 * (-p0 + p2);
    const double d = p1;
    *f = d + x * (c + x * (b + x * a));
    if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

inline void bicubic(const float* img, int w, int h, double r, double c, double* f, double* dfdr, double* dfdc)
{
    const int row = static_cast<int>(std::floor(r));
    const int col = static_cast<int>(std::floor(c));
    double fr[4], dfc[4];
    for (int i = 0; i < 4; ++i)
    {
        const int rr = std::min(std::max(0, row - 1 + i), h - 1);
        double p[4];
        for (int j = 0; j < 4; ++j)
        {
            const int cc = std::min(std::max(0, col - 1 + j), w - 1);
            p[j] = static_cast<double>(img[static_cast<size_t>(rr) * w + cc]);
        }
        cubic_hermite(p[0], p[1], p[2], p[3], c - col, &fr[i], &dfc[i]);
    }
    cubic_hermite(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
    double unused;
    cubic_hermite(dfc[0], dfc[1], dfc[2], dfc[3], r - row, dfdc, nullptr);
    (void)unused;
}

inline bool sample(const float* img, int w, int h, const double uv[2], double* out)
{
    double f, dr, dc;
    bicubic(img, w, h, uv[1], uv[0], &f, &dr, &dc);
    if (std::isfinite(f)) { *out = f; return true; }
    *out = 0.0; return false;
}
inline bool sample(const float* img, int w, int h, const Jet uv[2], Jet* out)
{
    double f, dr, dc;
    bicubic(img, w, h, uv[1].a, uv[0].a, &f, &dr, &dc);
    Jet l; l.a = f;
    for (int i = 0; i < kStride; ++i) l.v[i] = dr * uv[1].v[i] + dc * uv[0].v[i];
    if (finite_all(l)) { *out = l; return true; }
    *out = Jet(0.0); return false;
}

// ---------------------------------------------------------------------------
// E_g functor (ShadingCost::operator(), include/nv/refinement/shading_cost.h:85-198)
// ---------------------------------------------------------------------------
struct EgContext
{
    int coord[3];
    double voxel_size;   // float voxel size widened to double (Q15)
    double pyr_scale;
    int w, h;
    const float* lum;
    const double* sh;    // 9 constants
};

// sdf parameter index -> which entry of each point's quadruple (s, s+x, s+y, s+z)
// point 0 = v, 1 = v+x, 2 = v+y, 3 = v+z   (shading_cost.h:133-137)
static const int kQuad[4][4] = { {0, 6, 1, 4}, {6, 9, 7, 8}, {1, 7, 2, 3}, {4, 8, 3, 5} };
static const int kPointOffset[4][3] = { {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1} };
// albedo parameter of each point: params 10,11,12,13 = (0,0,0),(1,0,0),(0,1,0),(0,0,1)
static const int kPointAlbedo[4] = {0, 1, 2, 3};

template <class T>
inline void rotate_angle_axis(const T aa[3], const T p[3], T out[3])
{
    // ceres::AngleAxisRotatePoint
    const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (scalar(theta2) > std::numeric_limits<double>::epsilon())
    {
        const T theta = jsqrt(theta2);
        const T ct = jcos(theta);
        const T st = jsin(theta);
        const T ti = T(1.0) / theta;
        const T w[3] = {aa[0] * ti, aa[1] * ti, aa[2] * ti};
        const T wxp[3] = {w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]};
        const T tmp = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (T(1.0) - ct);
        out[0] = p[0] * ct + wxp[0] * st + w[0] * tmp;
        out[1] = p[1] * ct + wxp[1] * st + w[1] * tmp;
        out[2] = p[2] * ct + wxp[2] * st + w[2] * tmp;
    }
    else
    {
        const T wxp[3] = {aa[1] * p[2] - aa[2] * p[1], aa[2] * p[0] - aa[0] * p[2], aa[0] * p[1] - aa[1] * p[0]};
        out[0] = p[0] + wxp[0];
        out[1] = p[1] + wxp[1];
        out[2] = p[2] + wxp[2];
    }
}

// returns the residual (NV_INVALID_RESIDUAL == 0.0 for invalid rows)
template <class T>
T eg_functor(const EgContext& ctx, const T sdf[10], const T alb[4], const T pose[6], const T intr[4], const T dist[5])
{
    const T zero(0.0);
    T n[4][3];
    T uv[4][2];
    const T vs(ctx.voxel_size);
    const T ps(ctx.pyr_scale);
    const T fx = intr[0] * ps, fy = intr[1] * ps, cx = intr[2] * ps, cy = intr[3] * ps;
    bool all_in = true;
    for (int i = 0; i < 4; ++i)
    {
        const T s = sdf[kQuad[i][0]];
        // SDFOperators::computeNormal (forward differences, normalised iff length > 0)
        T g[3] = {sdf[kQuad[i][1]] - s, sdf[kQuad[i][2]] - s, sdf[kQuad[i][3]] - s};
        const T len = jsqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        if (scalar(len) > 0.0) { g[0] = g[0] / len; g[1] = g[1] / len; g[2] = g[2] / len; }
        n[i][0] = g[0]; n[i][1] = g[1]; n[i][2] = g[2];
        // voxelToWorld + voxelCenterToIso
        T X[3];
        for (int k = 0; k < 3; ++k)
            X[k] = T(static_cast<double>(ctx.coord[k] + kPointOffset[i][k])) * vs - n[i][k] * s;
        // transform: angle-axis rotation + translation
        T Y[3];
        rotate_angle_axis(pose, X, Y);
        Y[0] = Y[0] + pose[3]; Y[1] = Y[1] + pose[4]; Y[2] = Y[2] + pose[5];
        // CameraT::project
        T x = Y[0] / Y[2];
        T y = Y[1] / Y[2];
        const T r2 = x * x + y * y;
        const T r4 = r2 * r2;
        const T r6 = r4 * r2;
        const T dc = T(1.0) + dist[0] * r2 + dist[1] * r4 + dist[2] * r6;
        x = x * dc + T(2.0) * dist[3] * x * y + dist[4] * (r2 + T(2.0) * x * x);
        y = y * dc + T(2.0) * dist[4] * x * y + dist[3] * (r2 + T(2.0) * y * y);   // uses the distorted x (Q2)
        uv[i][0] = fx * x + cx;
        uv[i][1] = fy * y + cy;
        const double u = scalar(uv[i][0]), v = scalar(uv[i][1]);
        // NB: written as the negation of the reference's "outside" test so NaN behaves identically
        if (u < 0.0 || u > static_cast<double>(ctx.w - 1) || v < 0.0 || v > static_cast<double>(ctx.h - 1))
            all_in = false;
    }
    if (!all_in) return zero;

    T lum[4];
    bool ok = true;
    for (int i = 0; i < 4; ++i) ok = sample(ctx.lum, ctx.w, ctx.h, uv[i], &lum[i]) && ok;
    if (!ok) return zero;

    T shading[4];
    for (int i = 0; i < 4; ++i)
    {
        // Shading::shBasisFunctions (un-normalised, reference order; Q9)
        const T* m = n[i];
        T b[9];
        b[0] = T(1.0);
        b[1] = m[1];
        b[2] = m[2];
        b[3] = m[0];
        b[4] = m[0] * m[1];
        b[5] = m[1] * m[2];
        b[6] = (-(m[0] * m[0])) - (m[1] * m[1]) + T(2.0) * (m[2] * m[2]);
        b[7] = m[0] * m[2];
        b[8] = (m[0] * m[0]) - (m[1] * m[1]);
        T acc(0.0);
        for (int k = 0; k < 9; ++k) acc += T(ctx.sh[k]) * b[k];
        shading[i] = alb[kPointAlbedo[i]] * acc;
    }
    // Shading::computeShadingGradientDifference
    const T dx = (shading[1] - shading[0]) - (lum[1] - lum[0]);
    const T dy = (shading[2] - shading[0]) - (lum[2] - lum[0]);
    const T dz = (shading[3] - shading[0]) - (lum[3] - lum[0]);
    const T res = jsqrt(dx * dx + dy * dy + dz * dz);
    if (!finite_all(res)) return zero;
    return res;
}

// ---------------------------------------------------------------------------
// float helpers for observation selection (all arithmetic in float, no FMA
// contraction: this file is compiled with -ffp-contract=off)
// ---------------------------------------------------------------------------
inline float robust_kernel(float val, float thres = 2.0f)
{
    const float div = (1.0f + thres * val);
    return 1.0f / (div * div * div);
}

// math::poseVecAAToMat (Eigen::AngleAxisd(|w|, w/|w|).matrix()), then cast to float
inline void pose_to_mat_f(const double* pose, float R[9], float t[3])
{
    const double wx = pose[0], wy = pose[1], wz = pose[2];
    const double n2 = wx * wx + wy * wy + wz * wz;
    const double angle = std::sqrt(n2);
    double ax = wx, ay = wy, az = wz;
    if (n2 > 0.0) { ax = wx / angle; ay = wy / angle; az = wz / angle; }
    const double s = std::sin(angle), c = std::cos(angle);
    const double sx = s * ax, sy = s * ay, sz = s * az;
    const double c1x = (1.0 - c) * ax, c1y = (1.0 - c) * ay, c1z = (1.0 - c) * az;
    double M[9];
    double tmp;
    tmp = c1x * ay; M[1] = tmp - sz; M[3] = tmp + sz;
    tmp = c1x * az; M[2] = tmp + sy; M[6] = tmp - sy;
    tmp = c1y * az; M[5] = tmp - sx; M[7] = tmp + sx;
    M[0] = c1x * ax + c; M[4] = c1y * ay + c; M[8] = c1z * az + c;
    for (int i = 0; i < 9; ++i) R[i] = static_cast<float>(M[i]);
    for (int i = 0; i < 3; ++i) t[i] = static_cast<float>(pose[3 + i]);
}

// colour intensity (src/color_util.cpp:41-46)
inline float intensity_u8(const uint8_t* c) { return 0.299f * static_cast<float>(c[0]) + 0.587f * static_cast<float>(c[1]) + 0.114f * static_cast<float>(c[2]); }

// ---------------------------------------------------------------------------
// problem containers
// ---------------------------------------------------------------------------
struct Row
{
    int type;          // 0..3
    int voxel;         // owning voxel index
    int frame;         // E_g only, else -1
    int nb;            // E_a: neighbour voxel
    int ncols;
    int cols[I3D_EG_COLS];   // global unknown index (may point at fixed unknowns)
    double w_raw;      // residual weight before type normalisation
    double sdf0;       // E_s constant
};

struct Oracle
{
    // grid
    int64_t n = 0;
    std::vector<int32_t> xyz;
    std::vector<double> sdf0, sdf, albedo;
    std::vector<float> weight;
    std::vector<uint8_t> rgb;
    float voxel_size = 0.f;
    float truncation = 0.f;
    std::unordered_map<uint64_t, int32_t> index;
    // frames
    int F = 0, W = 0, H = 0;
    std::vector<float> lum, depth;
    std::vector<uint8_t> color;      // [F][H][W][3] interleaved B,G,R like the reference's cv::Mat (CV_8UC3)
    double pyr_scale = 1.0;
    // camera
    std::vector<double> poses;    // 6F
    double intr[4] = {0, 0, 0, 0};
    double dist[5] = {0, 0, 0, 0, 0};
    // lighting
    std::vector<double> sh;       // 9n
    int threads = 8;
    int parallel_cg = 0;

    // last-iteration artefacts (for parity tests)
    std::vector<Row> rows;
    std::vector<double> row_res;        // unweighted residual at the initial point
    std::vector<double> row_w;          // final weight (raw * type weight)
    std::vector<double> eg_jac;         // [n_eg rows][29] raw (unweighted, unscaled) d r / d theta
    std::vector<int32_t> eg_row_ids;    // indices into rows of E_g rows
    std::vector<int32_t> obs_frame;     // [n][K] selected frames (-1 none), descending priority
    std::vector<float> obs_weight;      // [n][K]
    std::vector<uint8_t> active;        // [n]
    std::vector<uint8_t> free_mask;     // [2n + Pc]
    std::vector<double> last_step;      // delta (unscaled) of the last evaluated trial, [2n+Pc]
    std::vector<double> col_scale;      // jacobi scaling
    // SVSH lighting artefacts
    std::vector<int32_t> sub_index;     // [S][3] integer subvolume indices, ascending (z, y, x)
    std::vector<double> sub_sh;         // [S][9]
    std::vector<uint8_t> has_sh;        // [n] 1 where computeVoxelShCoeffs produced a vector
    std::vector<int32_t> light_row_sub, light_row_voxel;   // SHDataCost rows of the last estimate
    std::vector<double> light_row_j, light_row_lum, light_row_w;
    std::vector<int32_t> light_pairs;   // [P][2] directed subvolume pairs
    std::string error;

    static uint64_t key(int x, int y, int z)
    {
        const uint64_t B = 1u << 20;
        return ((static_cast<uint64_t>(x + static_cast<int64_t>(B)) & 0x1FFFFF) << 42) |
               ((static_cast<uint64_t>(y + static_cast<int64_t>(B)) & 0x1FFFFF) << 21) |
               (static_cast<uint64_t>(z + static_cast<int64_t>(B)) & 0x1FFFFF);
    }
    int find(int x, int y, int z) const
    {
        auto it = index.find(key(x, y, z));
        return it == index.end() ? -1 : it->second;
    }
    bool valid_idx(int i) const { return i >= 0 && weight[i] > 0.0f; }

    int64_t num_unknowns() const { return 2 * n + 6 * static_cast<int64_t>(F) + 9; }
    int64_t col_pose(int f) const { return 2 * n + 6 * static_cast<int64_t>(f); }
    int64_t col_intr() const { return 2 * n + 6 * static_cast<int64_t>(F); }
    int64_t col_dist() const { return col_intr() + 4; }
};

// SDFOperators::computeSurfaceNormal (float; src/sdf/operators.cpp:58-77)
inline bool surface_normal_f(const Oracle& o, int v, float nrm[3])
{
    const int x = o.xyz[3 * v], y = o.xyz[3 * v + 1], z = o.xyz[3 * v + 2];
    const int ix = o.find(x + 1, y, z), iy = o.find(x, y + 1, z), iz = o.find(x, y, z + 1);
    nrm[0] = nrm[1] = nrm[2] = 0.0f;
    if (!o.valid_idx(v) || !o.valid_idx(ix) || !o.valid_idx(iy) || !o.valid_idx(iz)) return false;
    const float s0 = static_cast<float>(o.sdf[v]);
    float g0 = static_cast<float>(o.sdf[ix]) - s0;
    float g1 = static_cast<float>(o.sdf[iy]) - s0;
    float g2 = static_cast<float>(o.sdf[iz]) - s0;
    // Eigen: squaredNorm then sqrt; normalise iff norm != 0
    const float sq = g0 * g0 + g1 * g1 + g2 * g2;
    const float len = std::sqrt(sq);
    if (len != 0.0f) { g0 = g0 / len; g1 = g1 / len; g2 = g2 / len; }
    nrm[0] = g0; nrm[1] = g1; nrm[2] = g2;
    return !(g0 == 0.0f && g1 == 0.0f && g2 == 0.0f);
}

inline bool ring_valid(const Oracle& o, int v, int nb[6])
{
    const int x = o.xyz[3 * v], y = o.xyz[3 * v + 1], z = o.xyz[3 * v + 2];
    // order +x,-x,+y,-y,+z,-z (src/sdf/algorithms.cpp:75-91)
    nb[0] = o.find(x + 1, y, z); nb[1] = o.find(x - 1, y, z);
    nb[2] = o.find(x, y + 1, z); nb[3] = o.find(x, y - 1, z);
    nb[4] = o.find(x, y, z + 1); nb[5] = o.find(x, y, z - 1);
    bool ok = true;
    for (int i = 0; i < 6; ++i) if (!o.valid_idx(nb[i])) ok = false;
    return ok;
}

// SDFColorization::computeObservation weight for one frame (float pipeline)
inline float observation_weight(const Oracle& o, int v, const float nrm[3], const float R[9], const float t[3],
                                float fx, float fy, float cx, float cy, const float distf[5], bool dist_zero,
                                const float* depth, float occlusion, float* pix_out = nullptr)
{
    // voxelCenterToIso(grid, v, n): pt = float(coord)*voxel_size - n*float(sdf_refined)
    const float s = static_cast<float>(o.sdf[v]);
    float pt[3];
    for (int k = 0; k < 3; ++k) pt[k] = static_cast<float>(o.xyz[3 * v + k]) * o.voxel_size - nrm[k] * s;
    // pose_world_to_cam.topLeftCorner(3,3) * pt + t
    float q[3];
    for (int k = 0; k < 3; ++k) q[k] = ((R[3 * k] * pt[0] + R[3 * k + 1] * pt[1]) + R[3 * k + 2] * pt[2]) + t[k];
    // Camera::project (src/camera.cpp:124-154)
    float x = q[0] / q[2];
    float y = q[1] / q[2];
    if (!dist_zero)
    {
        const float r2 = x * x + y * y;
        const float r4 = r2 * r2;
        const float r6 = r4 * r2;
        const float dc = ((1.0f + distf[0] * r2) + distf[1] * r4) + distf[2] * r6;
        x = (x * dc + ((2.0f * distf[3]) * x) * y) + distf[4] * (r2 + (2.0f * x) * x);
        y = (y * dc + ((2.0f * distf[4]) * x) * y) + distf[3] * (r2 + (2.0f * y) * y);
    }
    const float pu = fx * x + cx;
    const float pv = fy * y + cy;
    if (pix_out) { pix_out[0] = pu; pix_out[1] = pv; }     // Camera::project's pt2f (sub-pixel position for the colour lookup)
    // static_cast<int>(p + 0.5f) — guard against UB for non-finite / huge values
    const float pu5 = pu + 0.5f, pv5 = pv + 0.5f;
    if (!(pu5 > -2147483000.0f && pu5 < 2147483000.0f && pv5 > -2147483000.0f && pv5 < 2147483000.0f)) return 0.0f;
    const int iu = static_cast<int>(pu5), iv = static_cast<int>(pv5);
    if (iu < 0 || iu >= o.W || iv < 0 || iv >= o.H) return 0.0f;
    const float d = depth[static_cast<size_t>(iv) * o.W + iu];
    // isVoxelVisible
    if (occlusion > 0.0f)
    {
        if (!(d > 0.0f)) return 0.0f;
        const float sd = d - q[2];
        if (!(std::fabs(sd) <= occlusion)) return 0.0f;
    }
    // computeWeight
    if (d <= 0.0f) return 0.0f;
    float nc[3];
    for (int k = 0; k < 3; ++k) nc[k] = (R[3 * k] * nrm[0] + R[3 * k + 1] * nrm[1]) + R[3 * k + 2] * nrm[2];
    float w_normal = 0.0f;
    if (!(nc[0] == 0.0f && nc[1] == 0.0f && nc[2] == 0.0f))
    {
        const float qn2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
        float vd[3] = {q[0], q[1], q[2]};
        if (qn2 > 0.0f) { const float ql = std::sqrt(qn2); vd[0] = q[0] / ql; vd[1] = q[1] / ql; vd[2] = q[2] / ql; }
        const float dt = (vd[0] * nc[0] + vd[1] * nc[1]) + vd[2] * nc[2];
        w_normal = 1.0f - std::fabs(dt);
        w_normal = std::max(std::min(w_normal, 1.0f), 0.0f);
        w_normal = std::max(robust_kernel(w_normal), 0.001f);
    }
    const float d_min = 0.01f, d_max = 5.0f;
    const float dw = std::max(std::min(d_max, d), d_min);
    const float depth_normalized = (dw - d_min) / (d_max - d_min);
    float w_depth = std::max(1.0f - depth_normalized, 1.0f);   // == 1 (Q1)
    w_depth = std::max(std::min(w_depth, 5.0f), 0.001f);
    return w_normal * w_depth;
}

// AlbedoRegularizer::create weight (src/refinement/albedo_regularizer.cpp:60-72); NaN/Inf => skip
inline bool albedo_pair_weight(const Oracle& o, int a, int b, double* w_out)
{
    const uint8_t* ca = &o.rgb[3 * a];
    const uint8_t* cb = &o.rgb[3 * b];
    const float la = intensity_u8(ca), lb = intensity_u8(cb);
    float dsq = 0.0f;
    float d[3];
    for (int k = 0; k < 3; ++k)
    {
        const float fa = static_cast<float>(ca[k]) * (1.0f / 255.0f);
        const float fb = static_cast<float>(cb[k]) * (1.0f / 255.0f);
        d[k] = fa / la - fb / lb;
    }
    dsq = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    float chroma = std::sqrt(dsq);
    chroma = std::max(1.0f - chroma, 0.01f);   // std::max(a,b): returns a when (a<b) is false => NaN stays NaN
    const double w = static_cast<double>(chroma) * 1.0;
    if (std::isnan(w) || std::isinf(w)) return false;
    *w_out = w;
    return true;
}

using Clock = std::chrono::steady_clock;
inline double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// ---------------------------------------------------------------------------
// residual evaluation at an arbitrary state
// ---------------------------------------------------------------------------
struct State
{
    const double* sdf; const double* albedo; const double* poses; const double* intr; const double* dist;
};

inline void eg_gather(const Oracle& o, const Row& r, const State& s, double sdf[10], double alb[4])
{
    for (int k = 0; k < 10; ++k) sdf[k] = s.sdf[r.cols[k]];
    for (int k = 0; k < 4; ++k) alb[k] = s.albedo[r.cols[10 + k] - o.n];
}

inline EgContext eg_context(const Oracle& o, const Row& r)
{
    EgContext c;
    c.coord[0] = o.xyz[3 * r.voxel]; c.coord[1] = o.xyz[3 * r.voxel + 1]; c.coord[2] = o.xyz[3 * r.voxel + 2];
    c.voxel_size = static_cast<double>(o.voxel_size);
    c.pyr_scale = o.pyr_scale;
    c.w = o.W; c.h = o.H;
    c.lum = o.lum.data() + static_cast<size_t>(r.frame) * o.W * o.H;
    c.sh = o.sh.data() + 9 * static_cast<size_t>(r.voxel);
    return c;
}

inline double eval_row(const Oracle& o, const Row& r, const State& s)
{
    switch (r.type)
    {
    case 0:
    {
        double sdf[10], alb[4];
        eg_gather(o, r, s, sdf, alb);
        const EgContext c = eg_context(o, r);
        return eg_functor<double>(c, sdf, alb, s.poses + 6 * r.frame, s.intr, s.dist);
    }
    case 1:
    {
        // computeLaplacian (include/nv/sdf/operators.h:90-109); params self,+x,-x,+y,-y,+z,-z
        const double c = s.sdf[r.cols[0]];
        const double dxx = s.sdf[r.cols[1]] + s.sdf[r.cols[2]] - 2.0 * c;
        const double dyy = s.sdf[r.cols[3]] + s.sdf[r.cols[4]] - 2.0 * c;
        const double dzz = s.sdf[r.cols[5]] + s.sdf[r.cols[6]] - 2.0 * c;
        return dxx + dyy + dzz;
    }
    case 2:
    {
        double res = s.sdf[r.cols[0]] - r.sdf0;
        if (res == 0.0) res = 0.0000001;   // Q3
        return res;
    }
    default:
        return s.albedo[r.cols[0] - o.n] - s.albedo[r.cols[1] - o.n];
    }
}

// raw Jacobian row (d residual / d params) for one row at state s; E_g via 8 passes of 4-lane Jets
inline void jac_row(const Oracle& o, const Row& r, const State& s, double* jac /* ncols */)
{
    switch (r.type)
    {
    case 0:
    {
        double sdf[10], alb[4];
        eg_gather(o, r, s, sdf, alb);
        const EgContext c = eg_context(o, r);
        const double* pose = s.poses + 6 * r.frame;
        for (int base = 0; base < I3D_EG_COLS; base += kStride)
        {
            Jet jsdf[10], jalb[4], jpose[6], jintr[4], jdist[5];
            for (int k = 0; k < 10; ++k) jsdf[k] = Jet(sdf[k]);
            for (int k = 0; k < 4; ++k) jalb[k] = Jet(alb[k]);
            for (int k = 0; k < 6; ++k) jpose[k] = Jet(pose[k]);
            for (int k = 0; k < 4; ++k) jintr[k] = Jet(s.intr[k]);
            for (int k = 0; k < 5; ++k) jdist[k] = Jet(s.dist[k]);
            for (int l = 0; l < kStride && base + l < I3D_EG_COLS; ++l)
            {
                const int p = base + l;
                if (p < 10) jsdf[p].v[l] = 1.0;
                else if (p < 14) jalb[p - 10].v[l] = 1.0;
                else if (p < 20) jpose[p - 14].v[l] = 1.0;
                else if (p < 24) jintr[p - 20].v[l] = 1.0;
                else jdist[p - 24].v[l] = 1.0;
            }
            const Jet res = eg_functor<Jet>(c, jsdf, jalb, jpose, jintr, jdist);
            for (int l = 0; l < kStride && base + l < I3D_EG_COLS; ++l) jac[base + l] = res.v[l];
        }
        break;
    }
    case 1:
        jac[0] = -6.0; for (int k = 1; k < 7; ++k) jac[k] = 1.0;
        break;
    case 2:
    {
        const double res = s.sdf[r.cols[0]] - r.sdf0;
        jac[0] = (res == 0.0) ? 0.0 : 1.0;   // T(1e-7) constant => zero derivative (Q3)
        break;
    }
    default:
        jac[0] = 1.0; jac[1] = -1.0;
    }
}

// ---------------------------------------------------------------------------
// sparse Jacobian in CSR (scaled by sqrt(w) and the Jacobi column scale; fixed
// columns dropped)
// ---------------------------------------------------------------------------
struct Csr
{
    std::vector<int64_t> ptr;
    std::vector<int64_t> col;
    std::vector<double> val;
    int64_t rows = 0, cols = 0;
    // optional transpose for parallel J^T products
    std::vector<int64_t> tptr, trow;
    std::vector<double> tval;
};

void csr_right(const Csr& A, const double* x, double* y, int threads)   // y = A x
{
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
    for (int64_t i = 0; i < A.rows; ++i)
    {
        double acc = 0.0;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) acc += A.val[k] * x[A.col[k]];
        y[i] = acc;
    }
}

void csr_left(const Csr& A, const double* y, double* x, int threads)    // x = A^T y
{
    if (threads > 1 && !A.tptr.empty())
    {
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int64_t j = 0; j < A.cols; ++j)
        {
            double acc = 0.0;
            for (int64_t k = A.tptr[j]; k < A.tptr[j + 1]; ++k) acc += A.tval[k] * y[A.trow[k]];
            x[j] = acc;
        }
        return;
    }
    std::fill(x, x + A.cols, 0.0);
    for (int64_t i = 0; i < A.rows; ++i)
    {
        const double yi = y[i];
        if (yi == 0.0) continue;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) x[A.col[k]] += A.val[k] * yi;
    }
}

void csr_build_transpose(Csr& A)
{
    A.tptr.assign(A.cols + 1, 0);
    for (int64_t c : A.col) A.tptr[c + 1]++;
    for (int64_t j = 0; j < A.cols; ++j) A.tptr[j + 1] += A.tptr[j];
    A.trow.resize(A.col.size()); A.tval.resize(A.col.size());
    std::vector<int64_t> fill(A.tptr.begin(), A.tptr.end() - 1);
    for (int64_t i = 0; i < A.rows; ++i)
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        {
            const int64_t p = fill[A.col[k]]++;
            A.trow[p] = i; A.tval[p] = A.val[k];
        }
}

inline double dot(const std::vector<double>& a, const std::vector<double>& b)
{
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

// dense SPD inverse via Cholesky (BlockRandomAccessDiagonalMatrix::Invert: llt().solve(I))
bool spd_inverse(int m, const double* A, double* inv)
{
    std::vector<double> L(m * m, 0.0);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j <= i; ++j)
        {
            double s = A[i * m + j];
            for (int k = 0; k < j; ++k) s -= L[i * m + k] * L[j * m + k];
            if (i == j) { if (!(s > 0.0)) return false; L[i * m + i] = std::sqrt(s); }
            else L[i * m + j] = s / L[j * m + j];
        }
    for (int c = 0; c < m; ++c)
    {
        std::vector<double> y(m, 0.0), x(m, 0.0);
        for (int i = 0; i < m; ++i)
        {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= L[i * m + k] * y[k];
            y[i] = s / L[i * m + i];
        }
        for (int i = m - 1; i >= 0; --i)
        {
            double s = y[i];
            for (int k = i + 1; k < m; ++k) s -= L[k * m + i] * x[k];
            x[i] = s / L[i * m + i];
        }
        for (int i = 0; i < m; ++i) inv[i * m + c] = x[i];
    }
    return true;
}

// block-Jacobi preconditioner: 1x1 for voxel unknowns, 6x6 per pose, 4x4, 5x5
struct BlockJacobi
{
    int64_t nvox2 = 0;      // 2n scalar blocks
    int F = 0;
    std::vector<double> dinv;      // [2n] inverse of scalar blocks
    std::vector<double> pose_inv;  // [F][36]
    double intr_inv[16];
    double dist_inv[25];

    void apply(const double* r, double* z) const
    {
        for (int64_t j = 0; j < nvox2; ++j) z[j] = dinv[j] * r[j];
        for (int f = 0; f < F; ++f)
        {
            const double* M = &pose_inv[36 * f];
            const double* rr = r + nvox2 + 6 * f;
            double* zz = z + nvox2 + 6 * f;
            for (int i = 0; i < 6; ++i) { double s = 0.0; for (int k = 0; k < 6; ++k) s += M[6 * i + k] * rr[k]; zz[i] = s; }
        }
        {
            const double* rr = r + nvox2 + 6 * F; double* zz = z + nvox2 + 6 * F;
            for (int i = 0; i < 4; ++i) { double s = 0.0; for (int k = 0; k < 4; ++k) s += intr_inv[4 * i + k] * rr[k]; zz[i] = s; }
            rr += 4; zz += 4;
            for (int i = 0; i < 5; ++i) { double s = 0.0; for (int k = 0; k < 5; ++k) s += dist_inv[5 * i + k] * rr[k]; zz[i] = s; }
        }
    }
};

} // namespace

// ===========================================================================
// one Gauss-Newton (outer) iteration
// ===========================================================================
static int oracle_gn_iteration_impl(Oracle& o, const I3DParams& P, I3DIterInfo& info)
{
    std::memset(&info, 0, sizeof(info));
    const int64_t n = o.n;
    const int F = o.F;
    if (n <= 0 || F <= 0 || o.sh.size() != static_cast<size_t>(9 * n)) { o.error = "oracle: grid/frames/sh not set"; return 1; }
    int K = P.num_observations;
    if (K <= 0 || K > F) K = F;     // filter(): n == 0 or n >= num_obs keeps everything
    if (K > I3D_MAX_OBS) { o.error = "oracle: num_observations > I3D_MAX_OBS"; return 1; }
    const int Kslots = K;
    info.num_voxels = n;

    auto t_add = Clock::now();
    // ---- camera for observation selection: intrinsics * pyr_scale cast to float (Q13) ----
    const float fx = static_cast<float>(o.intr[0] * o.pyr_scale), fy = static_cast<float>(o.intr[1] * o.pyr_scale);
    const float cxf = static_cast<float>(o.intr[2] * o.pyr_scale), cyf = static_cast<float>(o.intr[3] * o.pyr_scale);
    float distf[5]; bool dist_zero = true;
    for (int k = 0; k < 5; ++k) { distf[k] = static_cast<float>(o.dist[k]); if (distf[k] != 0.0f) dist_zero = false; }
    std::vector<float> Rf(9 * static_cast<size_t>(F)), tf(3 * static_cast<size_t>(F));
    for (int f = 0; f < F; ++f) pose_to_mat_f(&o.poses[6 * f], &Rf[9 * f], &tf[3 * f]);

    const State x0{o.sdf.data(), o.albedo.data(), o.poses.data(), o.intr, o.dist};

    o.rows.clear(); o.row_res.clear(); o.row_w.clear(); o.eg_row_ids.clear();
    o.obs_frame.assign(static_cast<size_t>(n) * Kslots, -1);
    o.obs_weight.assign(static_cast<size_t>(n) * Kslots, 0.0f);
    o.active.assign(n, 0);
    std::vector<uint8_t> ringok(n, 0);

    // ---- pass A (parallelisable, order-independent): activity, observation selection ----
    // The reference does this inside its serial per-voxel loop; results do not depend on order.
#pragma omp parallel for num_threads(o.threads) schedule(dynamic, 256)
    for (int64_t v = 0; v < n; ++v)
    {
        int nb[6];
        ringok[v] = ring_valid(o, static_cast<int>(v), nb) ? 1 : 0;
        if (!o.valid_idx(static_cast<int>(v))) continue;
        if (std::fabs(o.sdf[v]) > P.thres_shell) continue;
        float nrm[3];
        if (!surface_normal_f(o, static_cast<int>(v), nrm)) continue;
        o.active[v] = 1;
        // collectObservations + filter (top-K by (weight, frame))
        std::vector<std::pair<float, int>> obs(F);
        for (int f = 0; f < F; ++f)
        {
            const float w = observation_weight(o, static_cast<int>(v), nrm, &Rf[9 * f], &tf[3 * f], fx, fy, cxf, cyf, distf, dist_zero,
                                               o.depth.data() + static_cast<size_t>(f) * o.W * o.H, P.occlusion_distance);
            obs[f] = std::make_pair(w, f);
        }
        std::sort(obs.begin(), obs.end());   // ascending (weight, frame); canonical tie-break
        for (int k = 0; k < Kslots; ++k)
        {
            const auto& e = obs[F - 1 - k];
            if (e.first > 0.0f) { o.obs_frame[v * Kslots + k] = e.second; o.obs_weight[v * Kslots + k] = e.first; }
        }
    }

    // ---- pass B (serial, in voxel order like optimizer.cpp:149-156): residual creation ----
    for (int64_t v = 0; v < n; ++v)
    {
        if (!o.active[v]) continue;
        info.num_active++;
        const int x = o.xyz[3 * v], y = o.xyz[3 * v + 1], z = o.xyz[3 * v + 2];
        const double weight_sdf = [&] {
            const double trunc = static_cast<double>(o.truncation);
            const double a = std::min(std::fabs(o.sdf[v]), trunc) / trunc;
            return std::min(std::max(1.0 - a, 0.01), 1.0);
        }();
        // E_g: ShadingCost::create per selected observation, ascending weight like the sorted vector
        const int i_y1 = o.find(x, y + 1, z), i_y2 = o.find(x, y + 2, z), i_y1z1 = o.find(x, y + 1, z + 1);
        const int i_z1 = o.find(x, y, z + 1), i_z2 = o.find(x, y, z + 2), i_x1 = o.find(x + 1, y, z);
        const int i_x1y1 = o.find(x + 1, y + 1, z), i_x1z1 = o.find(x + 1, y, z + 1), i_x2 = o.find(x + 2, y, z);
        const bool stencil_ok = i_x2 >= 0 && i_y2 >= 0 && i_z2 >= 0 && i_y1z1 >= 0 && i_x1y1 >= 0 && i_x1z1 >= 0;
        if (stencil_ok)
        {
            for (int k = Kslots - 1; k >= 0; --k)
            {
                const int f = o.obs_frame[v * Kslots + k];
                if (f < 0) continue;
                Row r; r.type = 0; r.voxel = static_cast<int>(v); r.frame = f; r.nb = -1; r.ncols = I3D_EG_COLS; r.sdf0 = 0.0;
                const int sidx[10] = {static_cast<int>(v), i_y1, i_y2, i_y1z1, i_z1, i_z2, i_x1, i_x1y1, i_x1z1, i_x2};
                for (int c = 0; c < 10; ++c) r.cols[c] = sidx[c];
                const int aidx[4] = {static_cast<int>(v), i_x1, i_y1, i_z1};
                for (int c = 0; c < 4; ++c) r.cols[10 + c] = static_cast<int>(n) + aidx[c];
                for (int c = 0; c < 6; ++c) r.cols[14 + c] = static_cast<int>(o.col_pose(f) + c);
                for (int c = 0; c < 4; ++c) r.cols[20 + c] = static_cast<int>(o.col_intr() + c);
                for (int c = 0; c < 5; ++c) r.cols[24 + c] = static_cast<int>(o.col_dist() + c);
                const double res = eval_row(o, r, x0);
                if (res == 0.0) continue;      // NV_INVALID_RESIDUAL: dropped at creation (Q4)
                r.w_raw = static_cast<double>(o.obs_weight[v * Kslots + k]) * weight_sdf;
                o.eg_row_ids.push_back(static_cast<int32_t>(o.rows.size()));
                o.rows.push_back(r); o.row_res.push_back(res);
            }
        }
        int nb[6];
        const bool rv = ring_valid(o, static_cast<int>(v), nb);
        if (P.use_er && rv)
        {
            Row r; r.type = 1; r.voxel = static_cast<int>(v); r.frame = -1; r.nb = -1; r.ncols = 7; r.w_raw = 1.0; r.sdf0 = 0.0;
            r.cols[0] = static_cast<int>(v); for (int c = 0; c < 6; ++c) r.cols[1 + c] = nb[c];
            o.rows.push_back(r); o.row_res.push_back(eval_row(o, r, x0));
        }
        if (P.use_es)
        {
            Row r; r.type = 2; r.voxel = static_cast<int>(v); r.frame = -1; r.nb = -1; r.ncols = 1; r.w_raw = 1.0; r.sdf0 = o.sdf0[v];
            r.cols[0] = static_cast<int>(v);
            o.rows.push_back(r); o.row_res.push_back(eval_row(o, r, x0));
        }
        if (P.use_ea && rv)
        {
            for (int c = 0; c < 6; ++c)
            {
                // voxels_added.find(nb): neighbour is an active voxel visited earlier (Q6)
                if (o.active[nb[c]] && nb[c] < v) continue;
                double w;
                if (!albedo_pair_weight(o, static_cast<int>(v), nb[c], &w)) continue;
                if (w == 0.0) continue;
                Row r; r.type = 3; r.voxel = static_cast<int>(v); r.frame = -1; r.nb = nb[c]; r.ncols = 2; r.w_raw = w; r.sdf0 = 0.0;
                r.cols[0] = static_cast<int>(n + v); r.cols[1] = static_cast<int>(n + nb[c]);
                o.rows.push_back(r); o.row_res.push_back(eval_row(o, r, x0));
            }
        }
    }
    info.time_add = seconds_since(t_add);
    if (info.num_active == 0) { info.termination = 4; return 0; }

    // ---- buildProblem: per-type weight normalisation (nls_solver.cpp:379-394), masks ----
    auto t_build = Clock::now();
    const int64_t nrows = static_cast<int64_t>(o.rows.size());
    double sum_w[4] = {0, 0, 0, 0};
    for (const Row& r : o.rows) { sum_w[r.type] += r.w_raw; info.type_residuals[r.type]++; }
    double type_w[4];
    for (int t = 0; t < 4; ++t)
    {
        type_w[t] = (sum_w[t] != 0.0) ? (P.lambda[t] / sum_w[t]) * 1000.0 : 0.0;
        info.type_sum_weights[t] = sum_w[t]; info.type_weights[t] = type_w[t];
    }
    o.row_w.resize(nrows);
    for (int64_t i = 0; i < nrows; ++i) o.row_w[i] = o.rows[i].w_raw * type_w[o.rows[i].type];

    const int64_t U = o.num_unknowns();
    o.free_mask.assign(U, 0);
    for (int64_t v = 0; v < n; ++v)
    {
        // fixVoxelParams (optimizer.cpp:312-361)
        bool fix = !o.valid_idx(static_cast<int>(v)) || std::fabs(o.sdf[v]) > P.thres_shell || !ringok[v];
        if (!fix) { o.free_mask[v] = 1; info.num_free_sdf++; }
        if (!fix && !P.fix_all_albedo) { o.free_mask[n + v] = 1; info.num_free_albedo++; }
    }
    if (!P.fix_poses) for (int64_t j = o.col_pose(0); j < o.col_intr(); ++j) o.free_mask[j] = 1;
    if (!P.fix_intrinsics) for (int k = 0; k < 4; ++k) o.free_mask[o.col_intr() + k] = 1;
    if (!P.fix_distortion) for (int k = 0; k < 5; ++k) o.free_mask[o.col_dist() + k] = 1;
    info.time_build = seconds_since(t_build);

    // ---- solve: Ceres TrustRegionMinimizer + LM + CGNR restated ----
    auto t_solve = Clock::now();
    // residuals f = sqrt(w) r and raw Jacobian rows at x0 (8 threads like options.num_threads = 8)
    std::vector<double> fvec(nrows);
    double cost0 = 0.0;
    for (int64_t i = 0; i < nrows; ++i)
    {
        fvec[i] = std::sqrt(o.row_w[i]) * o.row_res[i];
        const double c = 0.5 * o.row_w[i] * o.row_res[i] * o.row_res[i];
        cost0 += c; info.type_costs[o.rows[i].type] += c;
    }
    info.cost_initial = cost0; info.cost_final = cost0;
    o.last_step.assign(U, 0.0);
    if (P.build_only)
    {
        // still expose raw E_g Jacobians for parity tests
        o.eg_jac.assign(o.eg_row_ids.size() * I3D_EG_COLS, 0.0);
#pragma omp parallel for num_threads(o.threads) schedule(dynamic, 64)
        for (int64_t e = 0; e < static_cast<int64_t>(o.eg_row_ids.size()); ++e)
            jac_row(o, o.rows[o.eg_row_ids[e]], x0, &o.eg_jac[e * I3D_EG_COLS]);
        info.termination = 4;
        info.time_solve = seconds_since(t_solve);
        return 0;
    }

    Csr A; A.rows = nrows; A.cols = U; A.ptr.assign(nrows + 1, 0);
    {
        // per-row raw Jacobians
        std::vector<double> raw(static_cast<size_t>(nrows) * I3D_EG_COLS, 0.0);
#pragma omp parallel for num_threads(o.threads) schedule(dynamic, 64)
        for (int64_t i = 0; i < nrows; ++i) jac_row(o, o.rows[i], x0, &raw[i * I3D_EG_COLS]);
        o.eg_jac.assign(o.eg_row_ids.size() * I3D_EG_COLS, 0.0);
        for (size_t e = 0; e < o.eg_row_ids.size(); ++e)
            std::memcpy(&o.eg_jac[e * I3D_EG_COLS], &raw[static_cast<size_t>(o.eg_row_ids[e]) * I3D_EG_COLS], sizeof(double) * I3D_EG_COLS);
        for (int64_t i = 0; i < nrows; ++i)
        {
            int cnt = 0;
            for (int c = 0; c < o.rows[i].ncols; ++c) if (o.free_mask[o.rows[i].cols[c]]) cnt++;
            A.ptr[i + 1] = A.ptr[i] + cnt;
        }
        A.col.resize(A.ptr[nrows]); A.val.resize(A.ptr[nrows]);
        for (int64_t i = 0; i < nrows; ++i)
        {
            int64_t p = A.ptr[i];
            const double sw = std::sqrt(o.row_w[i]);
            for (int c = 0; c < o.rows[i].ncols; ++c)
            {
                const int64_t j = o.rows[i].cols[c];
                if (!o.free_mask[j]) continue;
                A.col[p] = j; A.val[p] = sw * raw[i * I3D_EG_COLS + c]; ++p;
            }
        }
    }
    // non-finite Jacobian/residual => Ceres evaluation failure
    for (double v : A.val) if (!std::isfinite(v)) { info.termination = 3; o.error = "oracle: non-finite jacobian"; info.time_solve = seconds_since(t_solve); return 0; }

    // jacobi scaling: 1/(1+sqrt(colnorm^2)), computed once at the initial point
    std::vector<double> colsq(U, 0.0);
    for (size_t k = 0; k < A.val.size(); ++k) colsq[A.col[k]] += A.val[k] * A.val[k];
    o.col_scale.assign(U, 0.0);
    for (int64_t j = 0; j < U; ++j) o.col_scale[j] = 1.0 / (1.0 + std::sqrt(colsq[j]));
    for (size_t k = 0; k < A.val.size(); ++k) A.val[k] *= o.col_scale[A.col[k]];
    if (o.parallel_cg) csr_build_transpose(A);
    const int cg_threads = o.parallel_cg ? o.threads : 1;

    // gradient tolerance check at iteration 0 (unscaled gradient max-norm)
    {
        std::vector<double> g(U, 0.0);
        csr_left(A, fvec.data(), g.data(), cg_threads);
        double gmax = 0.0;
        for (int64_t j = 0; j < U; ++j) if (o.free_mask[j]) gmax = std::max(gmax, std::fabs(g[j] / o.col_scale[j]));
        if (gmax <= P.gradient_tolerance) { info.termination = 1; info.time_solve = seconds_since(t_solve); return 0; }
    }
    int64_t nparams = 0; double xnorm2 = 0.0;
    for (int64_t j = 0; j < U; ++j) if (o.free_mask[j] && colsq[j] > 0.0)
    {
        nparams++;
        double xv;
        if (j < n) xv = o.sdf[j]; else if (j < 2 * n) xv = o.albedo[j - n];
        else if (j < o.col_intr()) xv = o.poses[j - 2 * n]; else if (j < o.col_dist()) xv = o.intr[j - o.col_intr()]; else xv = o.dist[j - o.col_dist()];
        xnorm2 += xv * xv;
    }
    info.num_parameters = nparams;
    const double x_norm = std::sqrt(xnorm2);

    double radius = P.initial_trust_region_radius;
    double decrease_factor = 2.0;
    bool reuse_diagonal = false;
    std::vector<double> diag(U, 0.0), D(U, 0.0);
    int invalid_steps = 0;
    info.termination = 2;
    info.trust_region_radius = radius;

    std::vector<double> b(U), xs(U), r(U), z(U), p(U), q(U), tmp_rows(nrows), model(nrows);
    std::vector<double> c_sdf(n), c_alb(n), c_poses(6 * static_cast<size_t>(F));
    BlockJacobi M; M.nvox2 = 2 * n; M.F = F; M.dinv.resize(2 * n); M.pose_inv.resize(36 * static_cast<size_t>(F));

    auto apply_lhs = [&](const double* xin, double* yout) {
        // CgnrLinearOperator::RightMultiply: y = A^T (A x) + D^2 x
        csr_right(A, xin, tmp_rows.data(), cg_threads);
        csr_left(A, tmp_rows.data(), yout, cg_threads);
        for (int64_t j = 0; j < U; ++j) yout[j] += D[j] * D[j] * xin[j];
    };

    for (int it = 1; it <= P.lm_steps; ++it)
    {
        const int slot = std::min(it - 1, I3D_MAX_LM_STEPS - 1);
        info.lm_iterations = it;
        if (!reuse_diagonal)
        {
            std::fill(diag.begin(), diag.end(), 0.0);
            for (size_t k = 0; k < A.val.size(); ++k) diag[A.col[k]] += A.val[k] * A.val[k];
            for (int64_t j = 0; j < U; ++j) diag[j] = std::min(std::max(diag[j], P.min_lm_diagonal), P.max_lm_diagonal);
        }
        for (int64_t j = 0; j < U; ++j) D[j] = std::sqrt(diag[j] / radius);

        // ---- CGNR: (A^T A + D^2) y = A^T f, block-Jacobi preconditioned ----
        csr_left(A, fvec.data(), b.data(), cg_threads);
        {
            // BlockJacobiPreconditioner::Update: block diag of A^T A, + D^2, inverted
            for (int64_t j = 0; j < 2 * n; ++j) M.dinv[j] = 0.0;
            std::vector<double> pose_blk(36 * static_cast<size_t>(F), 0.0);
            double intr_blk[16] = {0}, dist_blk[25] = {0};
            for (int64_t i = 0; i < nrows; ++i)
            {
                const int64_t b0 = A.ptr[i], b1 = A.ptr[i + 1];
                for (int64_t k = b0; k < b1; ++k)
                {
                    const int64_t j = A.col[k];
                    if (j < 2 * n) { M.dinv[j] += A.val[k] * A.val[k]; continue; }
                    for (int64_t l = b0; l < b1; ++l)
                    {
                        const int64_t jl = A.col[l];
                        if (jl < 2 * n) continue;
                        if (j < o.col_intr())
                        {
                            const int64_t f = (j - 2 * n) / 6;
                            if (jl >= o.col_pose(static_cast<int>(f)) && jl < o.col_pose(static_cast<int>(f)) + 6)
                                pose_blk[36 * f + 6 * (j - o.col_pose(static_cast<int>(f))) + (jl - o.col_pose(static_cast<int>(f)))] += A.val[k] * A.val[l];
                        }
                        else if (j < o.col_dist())
                        {
                            if (jl >= o.col_intr() && jl < o.col_dist()) intr_blk[4 * (j - o.col_intr()) + (jl - o.col_intr())] += A.val[k] * A.val[l];
                        }
                        else if (jl >= o.col_dist()) dist_blk[5 * (j - o.col_dist()) + (jl - o.col_dist())] += A.val[k] * A.val[l];
                    }
                }
            }
            for (int64_t j = 0; j < 2 * n; ++j) M.dinv[j] = 1.0 / (M.dinv[j] + D[j] * D[j]);
            bool ok = true;
            for (int f = 0; f < F; ++f)
            {
                for (int d = 0; d < 6; ++d) pose_blk[36 * f + 7 * d] += D[o.col_pose(f) + d] * D[o.col_pose(f) + d];
                ok = spd_inverse(6, &pose_blk[36 * f], &M.pose_inv[36 * f]) && ok;
            }
            for (int d = 0; d < 4; ++d) intr_blk[5 * d] += D[o.col_intr() + d] * D[o.col_intr() + d];
            for (int d = 0; d < 5; ++d) dist_blk[6 * d] += D[o.col_dist() + d] * D[o.col_dist() + d];
            ok = spd_inverse(4, intr_blk, M.intr_inv) && ok;
            ok = spd_inverse(5, dist_blk, M.dist_inv) && ok;
            if (!ok) { info.termination = 3; o.error = "oracle: preconditioner block not SPD"; break; }
        }
        std::fill(xs.begin(), xs.end(), 0.0);
        int cg_it = 0;
        bool cg_failed = false;
        const double norm_b = std::sqrt(dot(b, b));
        if (norm_b != 0.0)
        {
            r = b;
            double rho = 1.0;
            double Q0 = 0.0;    // -x.(b + r) with x = 0
            const int max_it = P.forced_cg_iterations > 0 ? P.forced_cg_iterations : P.max_linear_solver_iterations;
            for (cg_it = 1;; ++cg_it)
            {
                M.apply(r.data(), z.data());
                const double last_rho = rho;
                rho = dot(r, z);
                if (rho == 0.0 || !std::isfinite(rho)) { cg_failed = true; break; }
                if (cg_it == 1) p = z;
                else
                {
                    const double beta = rho / last_rho;
                    if (beta == 0.0 || !std::isfinite(beta)) { cg_failed = true; break; }
                    for (int64_t j = 0; j < U; ++j) p[j] = z[j] + beta * p[j];
                }
                apply_lhs(p.data(), q.data());
                const double pq = dot(p, q);
                if (pq <= 0.0 || std::isinf(pq)) break;      // NO_CONVERGENCE: step still used
                const double alpha = rho / pq;
                if (std::isinf(alpha)) { cg_failed = true; break; }
                for (int64_t j = 0; j < U; ++j) xs[j] += alpha * p[j];
                if (cg_it % P.residual_reset_period == 0)
                {
                    apply_lhs(xs.data(), z.data());
                    for (int64_t j = 0; j < U; ++j) r[j] = b[j] - z[j];
                }
                else
                    for (int64_t j = 0; j < U; ++j) r[j] -= alpha * q[j];
                double Q1 = 0.0;
                for (int64_t j = 0; j < U; ++j) Q1 -= xs[j] * (b[j] + r[j]);
                const double zeta = cg_it * (Q1 - Q0) / Q1;
                if (P.forced_cg_iterations > 0) { if (cg_it >= max_it) break; }
                else
                {
                    if (zeta < P.eta && cg_it >= P.min_linear_solver_iterations) break;
                    if (cg_it >= max_it) break;
                }
                Q0 = Q1;
            }
        }
        info.cg_iterations[slot] = cg_it; info.cg_iterations_total += cg_it;
        bool step_valid = !cg_failed;
        for (int64_t j = 0; j < U && step_valid; ++j) if (!std::isfinite(xs[j])) step_valid = false;
        double model_cost_change = 0.0;
        if (step_valid)
        {
            for (int64_t j = 0; j < U; ++j) xs[j] = -xs[j];      // LM strategy negates the solution
            csr_right(A, xs.data(), model.data(), cg_threads);
            for (int64_t i = 0; i < nrows; ++i) model_cost_change -= model[i] * (fvec[i] + model[i] / 2.0);
            step_valid = model_cost_change > 0.0;
        }
        info.model_cost_change[slot] = model_cost_change;
        if (!step_valid)
        {
            // HandleInvalidStep / LevenbergMarquardtStrategy::StepIsInvalid
            if (++invalid_steps >= P.max_consecutive_invalid_steps) { info.termination = 3; break; }
            radius *= 0.5; reuse_diagonal = true; info.trust_region_radius = radius;
            if (radius <= P.min_trust_region_radius) { info.termination = 1; break; }
            continue;
        }
        invalid_steps = 0;
        // undo column scaling, candidate point
        double step_norm2 = 0.0;
        for (int64_t j = 0; j < U; ++j) { o.last_step[j] = xs[j] * o.col_scale[j]; if (!o.free_mask[j]) o.last_step[j] = 0.0; step_norm2 += o.last_step[j] * o.last_step[j]; }
        info.step_norm = std::sqrt(step_norm2);
        for (int64_t v = 0; v < n; ++v) { c_sdf[v] = o.sdf[v] + o.last_step[v]; c_alb[v] = o.albedo[v] + o.last_step[n + v]; }
        for (int64_t k = 0; k < 6 * static_cast<int64_t>(F); ++k) c_poses[k] = o.poses[k] + o.last_step[2 * n + k];
        double c_intr[4], c_dist[5];
        for (int k = 0; k < 4; ++k) c_intr[k] = o.intr[k] + o.last_step[o.col_intr() + k];
        for (int k = 0; k < 5; ++k) c_dist[k] = o.dist[k] + o.last_step[o.col_dist() + k];
        const State xc{c_sdf.data(), c_alb.data(), c_poses.data(), c_intr, c_dist};
        double cand = 0.0;
#pragma omp parallel for num_threads(o.threads) schedule(dynamic, 256) reduction(+ : cand)
        for (int64_t i = 0; i < nrows; ++i)
        {
            const double rr = eval_row(o, o.rows[i], xc);
            cand += 0.5 * o.row_w[i] * rr * rr;
        }
        info.candidate_cost[slot] = cand;
        // ParameterToleranceReached / FunctionToleranceReached (return without applying the step)
        if (info.step_norm <= P.parameter_tolerance * (x_norm + P.parameter_tolerance)) { info.termination = 1; break; }
        const double cost_change = cost0 - cand;
        if (std::fabs(cost_change) <= P.function_tolerance * cost0) { info.termination = 1; break; }
        const double rho_q = cost_change / model_cost_change;
        info.relative_decrease[slot] = rho_q;
        if (rho_q > P.min_relative_decrease)
        {
            // HandleSuccessfulStep; the reference's callback then terminates the solve
            o.sdf = c_sdf; o.albedo = c_alb; o.poses = c_poses;
            for (int k = 0; k < 4; ++k) o.intr[k] = c_intr[k];
            for (int k = 0; k < 5; ++k) o.dist[k] = c_dist[k];
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho_q - 1.0, 3));
            radius = std::min(P.max_trust_region_radius, radius);
            info.trust_region_radius = radius;
            info.cost_final = cand; info.step_accepted = 1; info.termination = 0;
            break;
        }
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        info.trust_region_radius = radius;
        if (radius <= P.min_trust_region_radius) { info.termination = 1; break; }
    }
    info.time_solve = seconds_since(t_solve);
    return 0;
}


// ===========================================================================
// SVSH lighting: LightingSVSH::estimate (src/lighting/lighting_svsh.cpp:166-346)
// and LightingSVSH::computeVoxelShCoeffs (:93-110)
// ===========================================================================
// Subvolumes::pointToIndex (src/lighting/subvolumes.cpp:262-295): floor(pt * (1.0f / size_)), float arithmetic
inline int sub_point_to_index(float pt, float inv_size) { return static_cast<int>(std::floor(pt * inv_size)); }

// Shading::shBasisFunctions<double> on the float normal cast to double (include/nv/shading.h:53-67, lighting_svsh.cpp:125-134)
inline void sh_basis_d(const float nf[3], double b[9])
{
    const double n0 = nf[0], n1 = nf[1], n2 = nf[2];
    b[0] = 1.0; b[1] = n1; b[2] = n2; b[3] = n0; b[4] = n0 * n1; b[5] = n1 * n2;
    b[6] = (-n0 * n0) - (n1 * n1) + 2.0 * (n2 * n2); b[7] = n0 * n2; b[8] = (n0 * n0) - (n1 * n1);
}

// ceres::Solve on a LINEAR least-squares problem f(x) = A x + f0 (rows already scaled by sqrt(loss weight)),
// x0 = 0: TrustRegionMinimizer + LevenbergMarquardtStrategy + CgnrSolver(JACOBI on `block`-sized parameter blocks),
// run until a Ceres termination criterion fires (no callback here, unlike NLSSolver::solve).
static void lm_linear(const Csr& Araw, const std::vector<double>& f0, int block, const I3DLightingParams& P, std::vector<double>& x,
                      I3DLightingInfo& info)
{
    const int64_t M = Araw.cols, R = Araw.rows;
    x.assign(M, 0.0);
    std::vector<double> f(f0), cand_f(R), cand_x(M), g(M), b(M), xs(M), r(M), z(M), p(M), q(M), tmp(R), model(R), delta(M);
    auto cost_of = [](const std::vector<double>& v) { double c = 0.0; for (double t : v) c += t * t; return 0.5 * c; };
    double cost = cost_of(f);
    info.cost_initial = cost; info.cost_final = cost;
    // Jacobi scaling, once, at the initial point
    Csr A = Araw;
    std::vector<double> colsq(M, 0.0), scale(M, 1.0);
    for (size_t k = 0; k < A.val.size(); ++k) colsq[A.col[k]] += A.val[k] * A.val[k];
    for (int64_t j = 0; j < M; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(colsq[j]));
    for (size_t k = 0; k < A.val.size(); ++k) A.val[k] *= scale[A.col[k]];
    auto gradient_max = [&]() { csr_left(A, f.data(), g.data(), 1); double m = 0.0; for (int64_t j = 0; j < M; ++j) m = std::max(m, std::fabs(g[j] / scale[j])); return m; };
    double gmax = gradient_max();
    double x_norm = 0.0;
    double radius = P.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    std::vector<double> diag(M, 0.0), D(M, 0.0);
    const int nb = static_cast<int>(M / block);
    std::vector<double> blk(static_cast<size_t>(nb) * block * block), blk_inv(blk.size());
    int invalid_steps = 0;
    info.termination = 1;
    auto apply_lhs = [&](const double* xin, double* yout) {
        csr_right(A, xin, tmp.data(), 1);
        csr_left(A, tmp.data(), yout, 1);
        for (int64_t j = 0; j < M; ++j) yout[j] += D[j] * D[j] * xin[j];
    };
    int it = 0;
    for (;;)
    {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        info.trust_region_radius = radius;
        if (it >= P.max_iterations) { info.termination = 1; break; }
        if (gmax <= P.gradient_tolerance) { info.termination = 0; break; }
        if (radius <= P.min_trust_region_radius) { info.termination = 0; break; }
        ++it; info.lm_iterations = it;
        if (!reuse_diagonal)
        {
            std::fill(diag.begin(), diag.end(), 0.0);
            for (size_t k = 0; k < A.val.size(); ++k) diag[A.col[k]] += A.val[k] * A.val[k];
            for (int64_t j = 0; j < M; ++j) diag[j] = std::min(std::max(diag[j], P.min_lm_diagonal), P.max_lm_diagonal);
        }
        for (int64_t j = 0; j < M; ++j) D[j] = std::sqrt(diag[j] / radius);
        reuse_diagonal = true;
        csr_left(A, f.data(), b.data(), 1);
        // BlockJacobiPreconditioner: block diagonal of A^T A, + D^2, inverted
        std::fill(blk.begin(), blk.end(), 0.0);
        for (int64_t i = 0; i < R; ++i)
            for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
                for (int64_t l = A.ptr[i]; l < A.ptr[i + 1]; ++l)
                    if (A.col[k] / block == A.col[l] / block)
                        blk[static_cast<size_t>(A.col[k] / block) * block * block + (A.col[k] % block) * block + (A.col[l] % block)] += A.val[k] * A.val[l];
        bool pre_ok = true;
        for (int s = 0; s < nb; ++s)
        {
            double* B = &blk[static_cast<size_t>(s) * block * block];
            for (int d = 0; d < block; ++d) B[d * block + d] += D[s * block + d] * D[s * block + d];
            pre_ok = spd_inverse(block, B, &blk_inv[static_cast<size_t>(s) * block * block]) && pre_ok;
        }
        if (!pre_ok) { info.termination = 2; break; }
        auto precond = [&](const double* rin, double* zout) {
            for (int s = 0; s < nb; ++s)
            {
                const double* Bi = &blk_inv[static_cast<size_t>(s) * block * block];
                for (int i = 0; i < block; ++i) { double acc = 0.0; for (int k = 0; k < block; ++k) acc += Bi[i * block + k] * rin[s * block + k]; zout[s * block + i] = acc; }
            }
        };
        std::fill(xs.begin(), xs.end(), 0.0);
        int cg_it = 0; bool cg_failed = false;
        const double norm_b = std::sqrt(dot(b, b));
        if (norm_b != 0.0)
        {
            r = b;
            double rho = 1.0, Q0 = 0.0;
            for (cg_it = 1;; ++cg_it)
            {
                precond(r.data(), z.data());
                const double last_rho = rho;
                rho = dot(r, z);
                if (rho == 0.0 || std::isinf(rho)) { cg_failed = true; break; }
                if (cg_it == 1) p = z;
                else
                {
                    const double beta = rho / last_rho;
                    if (beta == 0.0 || std::isinf(beta)) { cg_failed = true; break; }
                    for (int64_t j = 0; j < M; ++j) p[j] = z[j] + beta * p[j];
                }
                apply_lhs(p.data(), q.data());
                const double pq = dot(p, q);
                if (pq <= 0.0 || std::isinf(pq)) break;
                const double alpha = rho / pq;
                if (std::isinf(alpha)) { cg_failed = true; break; }
                for (int64_t j = 0; j < M; ++j) xs[j] += alpha * p[j];
                if (cg_it % P.residual_reset_period == 0)
                {
                    apply_lhs(xs.data(), z.data());
                    for (int64_t j = 0; j < M; ++j) r[j] = b[j] - z[j];
                }
                else
                    for (int64_t j = 0; j < M; ++j) r[j] -= alpha * q[j];
                double Q1 = 0.0;
                for (int64_t j = 0; j < M; ++j) Q1 -= xs[j] * (b[j] + r[j]);
                const double zeta = cg_it * (Q1 - Q0) / Q1;
                if (zeta < P.eta && cg_it >= P.min_linear_solver_iterations) break;
                Q0 = Q1;
                if (cg_it >= P.max_linear_solver_iterations) break;
            }
        }
        info.cg_iterations_total += cg_it;
        bool step_valid = !cg_failed;
        for (int64_t j = 0; j < M && step_valid; ++j) if (!std::isfinite(xs[j])) step_valid = false;
        double model_cost_change = 0.0;
        if (step_valid)
        {
            for (int64_t j = 0; j < M; ++j) xs[j] = -xs[j];
            csr_right(A, xs.data(), model.data(), 1);
            for (int64_t i = 0; i < R; ++i) model_cost_change -= model[i] * (f[i] + model[i] / 2.0);
            step_valid = model_cost_change > 0.0;
        }
        if (!step_valid)
        {
            if (++invalid_steps >= P.max_consecutive_invalid_steps) { info.termination = 2; break; }
            radius *= 0.5; reuse_diagonal = true;
            continue;
        }
        invalid_steps = 0;
        double step_norm2 = 0.0;
        for (int64_t j = 0; j < M; ++j) { delta[j] = xs[j] * scale[j]; cand_x[j] = x[j] + delta[j]; const double dj = x[j] - cand_x[j]; step_norm2 += dj * dj; }
        csr_right(Araw, cand_x.data(), cand_f.data(), 1);
        for (int64_t i = 0; i < R; ++i) cand_f[i] += f0[i];
        const double cand = cost_of(cand_f);
        if (std::sqrt(step_norm2) <= P.parameter_tolerance * (x_norm + P.parameter_tolerance)) { info.termination = 0; break; }
        const double cost_change = cost - cand;
        if (std::fabs(cost_change) <= P.function_tolerance * cost) { info.termination = 0; break; }
        const double rho_q = cost_change / model_cost_change;
        if (rho_q > P.min_relative_decrease)
        {
            x = cand_x; f = cand_f; cost = cand;
            x_norm = std::sqrt(dot(x, x));
            gmax = gradient_max();
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho_q - 1.0, 3));
            radius = std::min(P.max_trust_region_radius, radius);
            decrease_factor = 2.0; reuse_diagonal = false;
            info.num_successful_steps++;
        }
        else { radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true; }
    }
    info.trust_region_radius = radius;
    info.cost_final = cost;
    info.usable = info.termination != 2;
}

static int oracle_lighting_impl(Oracle& o, const I3DLightingParams& P, I3DLightingInfo& info)
{
    std::memset(&info, 0, sizeof(info));
    o.sub_index.clear(); o.sub_sh.clear();
    const int64_t n = o.n;
    o.sh.assign(9 * static_cast<size_t>(n), 0.0); o.has_sh.assign(n, 0);
    info.termination = 2;
    // LightingSVSH::estimate :170 (early outs) — a non-positive subvolume size makes the reference register the same
    // parameter block twice in one residual block (ceres::Problem aborts), so it is rejected here.
    if (n == 0 || !(P.thres_shell > 0.0)) return 0;
    if (!(P.subvolume_size > 0.0f)) { o.error = "oracle: subvolume_size must be > 0"; return 1; }
    const auto t0 = Clock::now();
    const float inv_size = 1.0f / P.subvolume_size;
    // Subvolumes::generate (src/lighting/subvolumes.cpp:211-239): one subvolume per occupied cube, all voxels of the hash count
    std::map<std::array<int, 3>, int> sub;      // key (z, y, x): canonical numbering
    std::vector<std::array<int, 3>> vsub(n);
    for (int64_t v = 0; v < n; ++v)
    {
        std::array<int, 3> k;
        for (int d = 0; d < 3; ++d) k[2 - d] = sub_point_to_index(static_cast<float>(o.xyz[3 * v + d]) * o.voxel_size, inv_size);
        vsub[v] = k; sub[k] = 0;
    }
    int S = 0;
    for (auto& kv : sub) { kv.second = S++; o.sub_index.push_back(kv.first[2]); o.sub_index.push_back(kv.first[1]); o.sub_index.push_back(kv.first[0]); }
    info.num_subvolumes = S;
    auto find_sub = [&](int x, int y, int z) { auto it = sub.find({z, y, x}); return it == sub.end() ? -1 : it->second; };

    // data rows (lighting_svsh.cpp:195-252)
    struct DataRow { int sub; int voxel; double j[9]; double lum; double w; };
    std::vector<DataRow> rows;
    for (int64_t v = 0; v < n; ++v)
    {
        if (!o.valid_idx(static_cast<int>(v))) continue;
        if (std::fabs(o.sdf[v]) > P.thres_shell) continue;
        float nf[3];
        if (!surface_normal_f(o, static_cast<int>(v), nf)) continue;
        if (std::isnan(nf[0]) || std::isnan(nf[1]) || std::isnan(nf[2])) continue;
        const double albedo = o.albedo[v];
        if (albedo == 0.0 || std::isnan(albedo)) continue;
        DataRow r;
        double b[9]; sh_basis_d(nf, b);
        r.sub = sub[vsub[v]]; r.voxel = static_cast<int>(v);
        r.lum = static_cast<double>(intensity_u8(&o.rgb[3 * v]) / 255.0f);
        r.w = 1.0;
        if (P.weighted)
        {
            // SDFOperators::sdfToWeight (src/sdf/operators.cpp:142-147)
            const double T = static_cast<double>(o.truncation);
            const double a = std::min(std::fabs(o.sdf[v]), T) / T;
            r.w = std::min(std::max(1.0 - a, 0.01), 1.0);
        }
        for (int k = 0; k < 9; ++k) r.j[k] = albedo * b[k];     // d/dsh of albedo * sum(sh_k b_k) - lum
        rows.push_back(r);
    }
    // smoothness pairs (:255-289): directed, ring order +x,-x,+y,-y,+z,-z
    std::vector<std::pair<int, int>> pairs;
    static const int ring[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int i = 0; i < S; ++i)
        for (int d = 0; d < 6; ++d)
        {
            const int j = find_sub(o.sub_index[3 * i] + ring[d][0], o.sub_index[3 * i + 1] + ring[d][1], o.sub_index[3 * i + 2] + ring[d][2]);
            if (j >= 0) pairs.emplace_back(i, j);
        }
    info.num_data_rows = static_cast<int64_t>(rows.size());
    info.num_reg_pairs = static_cast<int64_t>(pairs.size());
    o.light_row_sub.clear(); o.light_row_voxel.clear(); o.light_row_j.clear(); o.light_row_lum.clear(); o.light_row_w.clear(); o.light_pairs.clear();
    for (const DataRow& r : rows)
    {
        o.light_row_sub.push_back(r.sub); o.light_row_voxel.push_back(r.voxel); o.light_row_lum.push_back(r.lum); o.light_row_w.push_back(r.w);
        o.light_row_j.insert(o.light_row_j.end(), r.j, r.j + 9);
    }
    for (const auto& pr : pairs) { o.light_pairs.push_back(pr.first); o.light_pairs.push_back(pr.second); }
    double sum_w = 0.0;
    for (const DataRow& r : rows) sum_w += r.w;
    info.sum_data_weights = sum_w;
    const double data_loss = sum_w > 0.0 ? 1.0 / sum_w : 1.0;                                             // :298-301
    const double reg_loss = pairs.empty() ? 0.0 : P.lambda_reg / static_cast<double>(pairs.size());       // :314
    // explicit Jacobian (ScaledLoss(nullptr, a): row * sqrt(a)), residual at sh = 0
    Csr A; A.cols = 9 * static_cast<int64_t>(S); A.rows = static_cast<int64_t>(rows.size()) + 9 * static_cast<int64_t>(pairs.size());
    A.ptr.assign(A.rows + 1, 0);
    std::vector<double> f0(A.rows, 0.0);
    for (size_t i = 0; i < rows.size(); ++i)
    {
        const double sw = std::sqrt(data_loss * rows[i].w);
        for (int k = 0; k < 9; ++k) { A.col.push_back(9 * static_cast<int64_t>(rows[i].sub) + k); A.val.push_back(sw * rows[i].j[k]); }
        A.ptr[i + 1] = static_cast<int64_t>(A.col.size());
        f0[i] = sw * (0.0 - rows[i].lum);
    }
    {
        const double sw = std::sqrt(reg_loss);
        int64_t rr = static_cast<int64_t>(rows.size());
        for (const auto& pr : pairs)
            for (int k = 0; k < 9; ++k)
            {
                A.col.push_back(9 * static_cast<int64_t>(pr.first) + k); A.val.push_back(sw);
                A.col.push_back(9 * static_cast<int64_t>(pr.second) + k); A.val.push_back(-sw);
                A.ptr[++rr] = static_cast<int64_t>(A.col.size());
            }
    }
    info.time_accumulate = seconds_since(t0);
    const auto t1 = Clock::now();
    lm_linear(A, f0, 9, P, o.sub_sh, info);
    info.time_solve = seconds_since(t1);
    if (!info.usable) return 0;

    // computeVoxelShCoeffs (:93-110) -> Subvolumes::interpolate(linear) (src/lighting/subvolumes.cpp:164-205)
    const auto t2 = Clock::now();
    for (int64_t v = 0; v < n; ++v)
    {
        if (!o.valid_idx(static_cast<int>(v)) || std::fabs(o.sdf[v]) > P.thres_shell) continue;
        float pos[3]; int v0[3]; float wgt[3];
        for (int d = 0; d < 3; ++d)
        {
            pos[d] = static_cast<float>(o.xyz[3 * v + d]) * o.voxel_size * inv_size - 0.5f;      // pointToIndexCoord
            v0[d] = static_cast<int>(std::floor(pos[d]));
            wgt[d] = pos[d] - static_cast<float>(v0[d]);
        }
        // math::interpolationWeights (src/math.cpp:103-128): corner order and float weight products
        static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
        float w8[8]; int id8[8];
        for (int c = 0; c < 8; ++c)
        {
            const float wx = corner[c][0] ? wgt[0] : (1.0f - wgt[0]);
            const float wy = corner[c][1] ? wgt[1] : (1.0f - wgt[1]);
            const float wz = corner[c][2] ? wgt[2] : (1.0f - wgt[2]);
            w8[c] = wx * wy * wz;
            id8[c] = find_sub(v0[0] + corner[c][0], v0[1] + corner[c][1], v0[2] + corner[c][2]);
            if (id8[c] < 0) w8[c] = 0.0f;
        }
        // math::average (src/math.cpp:74-96)
        double avg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float sum_w8 = 0.0f;
        for (int c = 0; c < 8; ++c)
        {
            const float w = w8[c];
            if (w == 0.0f) continue;
            const double wd = static_cast<double>(w);
            for (int k = 0; k < 9; ++k)
            {
                const double t = wd * o.sub_sh[9 * static_cast<size_t>(id8[c]) + k];
                avg[k] = (sum_w8 == 0.0f) ? t : avg[k] + t;
            }
            sum_w8 += w;
        }
        if (sum_w8 != 0.0f) { const double inv = static_cast<double>(1.0f / sum_w8); for (int k = 0; k < 9; ++k) avg[k] *= inv; }
        for (int k = 0; k < 9; ++k) o.sh[9 * static_cast<size_t>(v) + k] = avg[k];
        o.has_sh[v] = 1;
    }
    info.time_interpolate = seconds_since(t2);
    return 0;
}


// ===========================================================================
// voxel recolouring: Intrinsic3D::recomputeColors -> SDFColorization::add (per frame) + compute
// (src/refinement/intrinsic3d.cpp:381-409, src/sdf/colorization.cpp:113-189, 318-370)
// ===========================================================================
// interpolate<unsigned char> (src/rgbd/processing.cpp:236-291): bilinear on an interleaved 8-bit image, out-of-image taps dropped
inline unsigned char interp_u8(const uint8_t* img, int w, int h, int nc, float x, float y, int channel)
{
    int x0 = static_cast<int>(std::floor(x)), y0 = static_cast<int>(std::floor(y));
    const int x1 = x0 + 1, y1 = y0 + 1;
    float x1w = x - static_cast<float>(x0), y1w = y - static_cast<float>(y0);
    float x0w = 1.0f - x1w, y0w = 1.0f - y1w;
    if (x0 < 0 || x0 >= w) x0w = 0.0f;
    if (x1 < 0 || x1 >= w) x1w = 0.0f;
    if (y0 < 0 || y0 >= h) y0w = 0.0f;
    if (y1 < 0 || y1 >= h) y1w = 0.0f;
    const float w00 = x0w * y0w, w10 = x1w * y0w, w01 = x0w * y1w, w11 = x1w * y1w;
    const float sum_w = ((w00 + w10) + w01) + w11;
    float sum = 0.0f;
    if (w00 > 0.0f) sum += static_cast<float>(img[(static_cast<size_t>(y0) * w + x0) * nc + channel]) * w00;
    if (w01 > 0.0f) sum += static_cast<float>(img[(static_cast<size_t>(y1) * w + x0) * nc + channel]) * w01;
    if (w10 > 0.0f) sum += static_cast<float>(img[(static_cast<size_t>(y0) * w + x1) * nc + channel]) * w10;
    if (w11 > 0.0f) sum += static_cast<float>(img[(static_cast<size_t>(y1) * w + x1) * nc + channel]) * w11;
    return sum_w > 0.0f ? static_cast<unsigned char>(sum / sum_w) : static_cast<unsigned char>(0);
}

// counts[0] = voxels recoloured, counts[1] = observations with weight > 0 (before the top-K filter)
static int oracle_recolor_impl(Oracle& o, float occlusion, int K, int64_t counts[2])
{
    const int64_t n = o.n;
    const int F = o.F;
    counts[0] = counts[1] = 0;
    if (n == 0 || F == 0) return 0;
    if (o.color.size() != static_cast<size_t>(F) * o.W * o.H * 3) { o.error = "oracle: colour frames missing"; return 1; }
    if (K < 0) { o.error = "oracle: bad max_num_observations"; return 1; }
    const float fx = static_cast<float>(o.intr[0] * o.pyr_scale), fy = static_cast<float>(o.intr[1] * o.pyr_scale);
    const float cxf = static_cast<float>(o.intr[2] * o.pyr_scale), cyf = static_cast<float>(o.intr[3] * o.pyr_scale);
    float distf[5]; bool dist_zero = true;
    for (int k = 0; k < 5; ++k) { distf[k] = static_cast<float>(o.dist[k]); if (distf[k] != 0.0f) dist_zero = false; }
    std::vector<float> Rf(9 * static_cast<size_t>(F)), tf(3 * static_cast<size_t>(F));
    for (int f = 0; f < F; ++f) pose_to_mat_f(&o.poses[6 * f], &Rf[9 * f], &tf[3 * f]);
    const size_t img = static_cast<size_t>(o.W) * o.H;
    std::vector<uint8_t> out(o.rgb);
    int64_t n_col = 0, n_obs = 0;
    // note: add() erodes the depth map (erodeDiscontinuities) but then passes the ORIGINAL depth to computeObservation
    // (colorization.cpp:126,146) - the eroded copy is dead, so it is not restated.
#pragma omp parallel for num_threads(o.threads) schedule(dynamic, 256) reduction(+ : n_col, n_obs)
    for (int64_t v = 0; v < n; ++v)
    {
        float nrm[3];
        if (!surface_normal_f(o, static_cast<int>(v), nrm)) continue;          // add(): normal.isZero() => no observation
        struct Obs { float w; int f; unsigned char c[3]; };
        std::vector<Obs> obs;
        for (int f = 0; f < F; ++f)
        {
            float pix[2];
            const float w = observation_weight(o, static_cast<int>(v), nrm, &Rf[9 * f], &tf[3 * f], fx, fy, cxf, cyf, distf, dist_zero,
                                               o.depth.data() + img * f, occlusion, pix);
            if (!(w > 0.0f)) continue;
            Obs ob; ob.w = w; ob.f = f;
            const uint8_t* cimg = o.color.data() + img * f * 3;
            ob.c[0] = interp_u8(cimg, o.W, o.H, 3, pix[0], pix[1], 2);          // interpolateRGB: r = channel 2 of the BGR image
            ob.c[1] = interp_u8(cimg, o.W, o.H, 3, pix[0], pix[1], 1);
            ob.c[2] = interp_u8(cimg, o.W, o.H, 3, pix[0], pix[1], 0);
            obs.push_back(ob);
        }
        if (obs.empty()) continue;                                               // compute(): colour unchanged
        n_obs += static_cast<int64_t>(obs.size());
        // filter(): sort ascending by weight (ties: frame id, the canonical order), zero all but the best K
        if (K > 0 && static_cast<size_t>(K) < obs.size())
        {
            std::sort(obs.begin(), obs.end(), [](const Obs& a, const Obs& b) { return a.w < b.w || (a.w == b.w && a.f < b.f); });
            for (size_t i = 0; i + K < obs.size(); ++i) obs[i].w = 0.0f;
        }
        // (when the filter does not run - K == 0 or K >= #observations - the list stays in frame order, and computeColor sums
        //  in that order; when it runs the sum goes in ascending weight order.  Both are kept: float sums are order-sensitive.)
        // computeColor (colorization.cpp:318-354)
        const float scale_color = 1.0f / 255.0f;
        float c[3] = {0.0f, 0.0f, 0.0f}, wsum = 0.0f;
        for (const Obs& ob : obs)
        {
            const float ws = ob.w * scale_color;
            c[0] += static_cast<float>(ob.c[0]) * ws; c[1] += static_cast<float>(ob.c[1]) * ws; c[2] += static_cast<float>(ob.c[2]) * ws;
            wsum = wsum + ob.w;
        }
        if (wsum > 0.0f) { const float s = 255.0f / wsum; c[0] *= s; c[1] *= s; c[2] *= s; }
        for (int k = 0; k < 3; ++k) out[3 * v + k] = static_cast<unsigned char>(c[k]);
        n_col++;
    }
    o.rgb.swap(out);
    counts[0] = n_col; counts[1] = n_obs;
    return 0;
}


// ===========================================================================
// grid-level transitions (src/sdf/algorithms.cpp): clearVoxelsOutsideThinShell, upsample<VoxelSBR>
// ===========================================================================
static void oracle_reindex(Oracle& o)
{
    o.n = static_cast<int64_t>(o.sdf.size());
    o.index.clear(); o.index.reserve(static_cast<size_t>(o.n) * 2);
    for (int64_t i = 0; i < o.n; ++i) o.index[Oracle::key(o.xyz[3 * i], o.xyz[3 * i + 1], o.xyz[3 * i + 2])] = static_cast<int32_t>(i);
    o.sh.clear(); o.has_sh.clear(); o.sub_index.clear(); o.sub_sh.clear();
}

// SDFAlgorithms::clearVoxelsOutsideThinShell (algorithms.cpp:368-458).  Survivors keep their relative order.
static int oracle_clear_shell_impl(Oracle& o, double thres_shell)
{
    const int64_t n = o.n;
    std::vector<uint8_t> keep(n, 0);
    static const int ring9[9][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}};
    for (int64_t v = 0; v < n; ++v)
    {
        if (!o.valid_idx(static_cast<int>(v)) || std::fabs(o.sdf[v]) > thres_shell) continue;
        keep[v] = 1;
        for (int k = 0; k < 9; ++k)
        {
            const int nb = o.find(o.xyz[3 * v] + ring9[k][0], o.xyz[3 * v + 1] + ring9[k][1], o.xyz[3 * v + 2] + ring9[k][2]);
            if (nb >= 0) keep[nb] = 1;
        }
    }
    std::vector<uint8_t> keep2(keep);
    for (int64_t v = 0; v < n; ++v)
    {
        if (keep[v]) continue;
        const bool negative = o.sdf[v] < 0.0;
        bool crossing = false;
        for (int dz = -2; dz <= 2 && !crossing; ++dz)
            for (int dy = -2; dy <= 2 && !crossing; ++dy)
                for (int dx = -2; dx <= 2 && !crossing; ++dx)
                {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    const int nb = o.find(o.xyz[3 * v] + dx, o.xyz[3 * v + 1] + dy, o.xyz[3 * v + 2] + dz);
                    if (nb < 0) continue;
                    if (negative ? (o.sdf[nb] >= 0.0) : (o.sdf[nb] < 0.0)) crossing = true;
                }
        if (crossing) keep2[v] = 1;
    }
    int64_t m = 0;
    for (int64_t v = 0; v < n; ++v)
    {
        if (!keep2[v]) continue;
        for (int k = 0; k < 3; ++k) { o.xyz[3 * m + k] = o.xyz[3 * v + k]; o.rgb[3 * m + k] = o.rgb[3 * v + k]; }
        o.sdf0[m] = o.sdf0[v]; o.sdf[m] = o.sdf[v]; o.albedo[m] = o.albedo[v]; o.weight[m] = o.weight[v];
        ++m;
    }
    if (m == 0) { o.error = "oracle: no voxel survives"; oracle_reindex(o); return 1; }
    o.xyz.resize(3 * m); o.rgb.resize(3 * m); o.sdf0.resize(m); o.sdf.resize(m); o.albedo.resize(m); o.weight.resize(m);
    oracle_reindex(o);
    return 0;
}

// SDFAlgorithms::upsample<VoxelSBR> with interpolate<VoxelSBR> (algorithms.cpp:118-235); children of voxel i at 8 i + (4 z + 2 y + x)
static int oracle_upsample_impl(Oracle& o)
{
    const int64_t n = o.n, m = 8 * n;
    std::vector<int32_t> xyz(3 * m);
    std::vector<double> sdf0(m), sdf(m), alb(m);
    std::vector<float> wgt(m);
    std::vector<uint8_t> rgb(3 * m);
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
    for (int64_t v = 0; v < n; ++v)
        for (int z = 0; z < 2; ++z)
            for (int y = 0; y < 2; ++y)
                for (int x = 0; x < 2; ++x)
                {
                    const int64_t c = 8 * v + (4 * z + 2 * y + x);
                    // pos = p + 0.5 (x, y, z); math::interpolationWeights: v0 = floor(pos) = p, fractional weights 0 or 0.5
                    const float pos[3] = {static_cast<float>(o.xyz[3 * v]) + static_cast<float>(x) * 0.5f, static_cast<float>(o.xyz[3 * v + 1]) + static_cast<float>(y) * 0.5f,
                                          static_cast<float>(o.xyz[3 * v + 2]) + static_cast<float>(z) * 0.5f};
                    int v0[3]; float t[3];
                    for (int d = 0; d < 3; ++d) { v0[d] = static_cast<int>(std::floor(pos[d])); t[d] = pos[d] - static_cast<float>(v0[d]); }
                    float a_sdf = 0.0f, a_w = 0.0f, a_alb = 0.0f, a_ref = 0.0f, a_c[3] = {0.0f, 0.0f, 0.0f}, sum_w = 0.0f;
                    int cnt_valid = 0;
                    for (int k = 0; k < 8; ++k)
                    {
                        const int nb = o.find(v0[0] + corner[k][0], v0[1] + corner[k][1], v0[2] + corner[k][2]);
                        if (!o.valid_idx(nb)) continue;
                        const float w = ((corner[k][0] ? t[0] : 1.0f - t[0]) * (corner[k][1] ? t[1] : 1.0f - t[1])) * (corner[k][2] ? t[2] : 1.0f - t[2]);
                        a_sdf += w * static_cast<float>(o.sdf0[nb]);
                        for (int d = 0; d < 3; ++d) a_c[d] += w * static_cast<float>(o.rgb[3 * nb + d]);
                        a_w += w * o.weight[nb];
                        a_alb += w * static_cast<float>(o.albedo[nb]);
                        a_ref += w * static_cast<float>(o.sdf[nb]);
                        sum_w += w;
                        ++cnt_valid;
                    }
                    if (sum_w > 0.0f) { a_sdf /= sum_w; a_w /= sum_w; a_alb /= sum_w; a_ref /= sum_w; for (int d = 0; d < 3; ++d) a_c[d] /= sum_w; }
                    if (cnt_valid <= 4) a_w = 0.0f;
                    xyz[3 * c] = 2 * o.xyz[3 * v] + x; xyz[3 * c + 1] = 2 * o.xyz[3 * v + 1] + y; xyz[3 * c + 2] = 2 * o.xyz[3 * v + 2] + z;
                    sdf0[c] = static_cast<double>(a_sdf); sdf[c] = static_cast<double>(a_ref); alb[c] = static_cast<double>(a_alb);
                    wgt[c] = std::max(a_w, 0.0f);
                    for (int d = 0; d < 3; ++d) rgb[3 * c + d] = static_cast<unsigned char>(static_cast<int>(a_c[d] + 0.5f));   // nv::round(Vec3f), include/nv/mat.h:90
                }
    o.xyz.swap(xyz); o.sdf0.swap(sdf0); o.sdf.swap(sdf); o.albedo.swap(alb); o.weight.swap(wgt); o.rgb.swap(rgb);
    o.voxel_size = o.voxel_size * 0.5f; o.truncation = o.voxel_size * 5.0f;
    oracle_reindex(o);
    return 0;
}

// ===========================================================================
// C API (ctypes-friendly)
// ===========================================================================
extern "C" {

void i3do_default_params(I3DParams* p)
{
    std::memset(p, 0, sizeof(*p));
    p->lambda[0] = 0.2; p->lambda[1] = 80.0; p->lambda[2] = 120.0; p->lambda[3] = 0.1;   // data/intrinsic3d.yml
    p->use_er = p->use_es = p->use_ea = 1;
    p->occlusion_distance = 0.02f; p->num_observations = 5; p->lm_steps = 50;
    p->initial_trust_region_radius = 1e4; p->max_trust_region_radius = 1e16; p->min_trust_region_radius = 1e-32;
    p->min_relative_decrease = 1e-3; p->min_lm_diagonal = 1e-6; p->max_lm_diagonal = 1e32; p->eta = 0.1;
    p->function_tolerance = 1e-6; p->gradient_tolerance = 1e-10; p->parameter_tolerance = 1e-8;
    p->max_linear_solver_iterations = 500; p->min_linear_solver_iterations = 0; p->residual_reset_period = 10;
    p->max_consecutive_invalid_steps = 5;
}

void i3do_default_lighting_params(I3DLightingParams* p)
{
    std::memset(p, 0, sizeof(*p));
    p->subvolume_size = 0.2f; p->weighted = 1; p->lambda_reg = 10.0; p->thres_shell = 0.0;     // include/nv/refinement/intrinsic3d.h:81-82
    p->max_iterations = 50; p->max_linear_solver_iterations = 500; p->min_linear_solver_iterations = 0; p->residual_reset_period = 10;
    p->max_consecutive_invalid_steps = 5;
    p->initial_trust_region_radius = 1e4; p->max_trust_region_radius = 1e16; p->min_trust_region_radius = 1e-32;
    p->min_relative_decrease = 1e-3; p->min_lm_diagonal = 1e-6; p->max_lm_diagonal = 1e32; p->eta = 0.1;
    p->function_tolerance = 1e-6; p->gradient_tolerance = 1e-10; p->parameter_tolerance = 1e-8;
}

int i3do_estimate_lighting(void* h, const I3DLightingParams* p, I3DLightingInfo* info)
{
    return oracle_lighting_impl(*static_cast<Oracle*>(h), *p, *info);
}

int64_t i3do_num_subvolumes(void* h) { return static_cast<int64_t>(static_cast<Oracle*>(h)->sub_index.size() / 3); }

int i3do_get_lighting(void* h, int32_t* sub_index3, double* sh9)
{
    auto* o = static_cast<Oracle*>(h);
    if (sub_index3) std::memcpy(sub_index3, o->sub_index.data(), sizeof(int32_t) * o->sub_index.size());
    if (sh9) std::memcpy(sh9, o->sub_sh.data(), sizeof(double) * o->sub_sh.size());
    return 0;
}

int64_t i3do_num_lighting_rows(void* h, int what) { auto* o = static_cast<Oracle*>(h); return what == 0 ? static_cast<int64_t>(o->light_row_sub.size()) : static_cast<int64_t>(o->light_pairs.size() / 2); }

int i3do_get_lighting_rows(void* h, int32_t* sub, int32_t* voxel, double* j9, double* lum, double* w, int32_t* pairs2)
{
    auto* o = static_cast<Oracle*>(h);
    const size_t m = o->light_row_sub.size();
    if (sub) std::memcpy(sub, o->light_row_sub.data(), sizeof(int32_t) * m);
    if (voxel) std::memcpy(voxel, o->light_row_voxel.data(), sizeof(int32_t) * m);
    if (j9) std::memcpy(j9, o->light_row_j.data(), sizeof(double) * 9 * m);
    if (lum) std::memcpy(lum, o->light_row_lum.data(), sizeof(double) * m);
    if (w) std::memcpy(w, o->light_row_w.data(), sizeof(double) * m);
    if (pairs2) std::memcpy(pairs2, o->light_pairs.data(), sizeof(int32_t) * o->light_pairs.size());
    return 0;
}

int i3do_get_voxel_sh(void* h, double* sh9n, uint8_t* has_sh)
{
    auto* o = static_cast<Oracle*>(h);
    if (sh9n) std::memcpy(sh9n, o->sh.data(), sizeof(double) * o->sh.size());
    if (has_sh) { if (o->has_sh.size() == static_cast<size_t>(o->n)) std::memcpy(has_sh, o->has_sh.data(), o->n); else std::memset(has_sh, 1, o->n); }
    return 0;
}

void* i3do_create() { return new Oracle(); }
void i3do_destroy(void* h) { delete static_cast<Oracle*>(h); }
const char* i3do_last_error(void* h) { return static_cast<Oracle*>(h)->error.c_str(); }
void i3do_set_threads(void* h, int threads, int parallel_cg) { auto* o = static_cast<Oracle*>(h); o->threads = std::max(1, threads); o->parallel_cg = parallel_cg; }

int i3do_set_grid(void* h, int64_t n, const int32_t* xyz, const double* sdf0, const double* sdf_refined, const double* albedo,
                  const float* weight, const uint8_t* rgb, float voxel_size)
{
    auto* o = static_cast<Oracle*>(h);
    o->n = n; o->voxel_size = voxel_size; o->truncation = voxel_size * 5.0f;   // src/sparse_voxel_grid.cpp:48
    o->xyz.assign(xyz, xyz + 3 * n); o->sdf0.assign(sdf0, sdf0 + n); o->sdf.assign(sdf_refined, sdf_refined + n);
    o->albedo.assign(albedo, albedo + n); o->weight.assign(weight, weight + n); o->rgb.assign(rgb, rgb + 3 * n);
    o->index.clear(); o->index.reserve(static_cast<size_t>(n) * 2);
    for (int64_t i = 0; i < n; ++i) o->index[Oracle::key(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])] = static_cast<int32_t>(i);
    if (static_cast<int64_t>(o->index.size()) != n) { o->error = "oracle: duplicate voxel coordinates"; return 1; }
    return 0;
}

int i3do_set_frames(void* h, int F, int W, int H, const float* lum, const float* depth, double pyr_scale)
{
    auto* o = static_cast<Oracle*>(h);
    o->F = F; o->W = W; o->H = H; o->pyr_scale = pyr_scale;
    const size_t cnt = static_cast<size_t>(F) * W * H;
    o->lum.assign(lum, lum + cnt); o->depth.assign(depth, depth + cnt);
    return 0;
}

int i3do_set_camera(void* h, const double* poses, const double* intr, const double* dist)
{
    auto* o = static_cast<Oracle*>(h);
    o->poses.assign(poses, poses + 6 * static_cast<size_t>(o->F));
    for (int k = 0; k < 4; ++k) o->intr[k] = intr[k];
    for (int k = 0; k < 5; ++k) o->dist[k] = dist[k];
    return 0;
}

int i3do_set_color_frames(void* h, const uint8_t* bgr)
{
    auto* o = static_cast<Oracle*>(h);
    if (o->F <= 0) { o->error = "oracle: set the frames first"; return 1; }
    o->color.assign(bgr, bgr + static_cast<size_t>(o->F) * o->W * o->H * 3);
    return 0;
}

int i3do_recompute_colors(void* h, float occlusion, int K, int64_t* counts2)
{
    return oracle_recolor_impl(*static_cast<Oracle*>(h), occlusion, K, counts2);
}

int i3do_get_colors(void* h, uint8_t* rgb3n) { auto* o = static_cast<Oracle*>(h); std::memcpy(rgb3n, o->rgb.data(), o->rgb.size()); return 0; }

int64_t i3do_num_voxels(void* h) { return static_cast<Oracle*>(h)->n; }
int i3do_clear_voxels_outside_thin_shell(void* h, double thres_shell) { return oracle_clear_shell_impl(*static_cast<Oracle*>(h), thres_shell); }
int i3do_upsample_grid(void* h) { return oracle_upsample_impl(*static_cast<Oracle*>(h)); }
int i3do_get_grid(void* h, int32_t* xyz, double* sdf0, double* sdf_refined, double* albedo, float* weight, uint8_t* rgb, float* voxel_size)
{
    auto* o = static_cast<Oracle*>(h);
    const size_t n = static_cast<size_t>(o->n);
    if (xyz) std::memcpy(xyz, o->xyz.data(), 3 * n * sizeof(int32_t));
    if (sdf0) std::memcpy(sdf0, o->sdf0.data(), n * sizeof(double));
    if (sdf_refined) std::memcpy(sdf_refined, o->sdf.data(), n * sizeof(double));
    if (albedo) std::memcpy(albedo, o->albedo.data(), n * sizeof(double));
    if (weight) std::memcpy(weight, o->weight.data(), n * sizeof(float));
    if (rgb) std::memcpy(rgb, o->rgb.data(), 3 * n);
    if (voxel_size) *voxel_size = o->voxel_size;
    return 0;
}

int i3do_set_sh(void* h, const double* sh) { auto* o = static_cast<Oracle*>(h); o->sh.assign(sh, sh + 9 * o->n); return 0; }

int i3do_gn_iteration(void* h, const I3DParams* p, I3DIterInfo* info)
{
    auto* o = static_cast<Oracle*>(h);
    return oracle_gn_iteration_impl(*o, *p, *info);
}

int i3do_get_state(void* h, double* sdf_refined, double* albedo, double* poses, double* intr, double* dist)
{
    auto* o = static_cast<Oracle*>(h);
    if (sdf_refined) std::memcpy(sdf_refined, o->sdf.data(), sizeof(double) * o->n);
    if (albedo) std::memcpy(albedo, o->albedo.data(), sizeof(double) * o->n);
    if (poses) std::memcpy(poses, o->poses.data(), sizeof(double) * 6 * o->F);
    if (intr) std::memcpy(intr, o->intr, sizeof(double) * 4);
    if (dist) std::memcpy(dist, o->dist, sizeof(double) * 5);
    return 0;
}

int64_t i3do_num_rows(void* h, int type)
{
    auto* o = static_cast<Oracle*>(h);
    if (type < 0) return static_cast<int64_t>(o->rows.size());
    int64_t c = 0; for (const Row& r : o->rows) if (r.type == type) c++; return c;
}

/* rows of one type, in creation order: voxel, frame (E_g) or neighbour (E_a) or -1, unweighted residual, final weight, raw weight */
int i3do_get_rows(void* h, int type, int32_t* voxel, int32_t* aux, double* residual, double* weight, double* raw_weight)
{
    auto* o = static_cast<Oracle*>(h);
    int64_t c = 0;
    for (size_t i = 0; i < o->rows.size(); ++i)
    {
        const Row& r = o->rows[i];
        if (r.type != type) continue;
        if (voxel) voxel[c] = r.voxel;
        if (aux) aux[c] = (type == 0) ? r.frame : r.nb;
        if (residual) residual[c] = o->row_res[i];
        if (weight) weight[c] = i < o->row_w.size() ? o->row_w[i] : 0.0;
        if (raw_weight) raw_weight[c] = r.w_raw;
        ++c;
    }
    return 0;
}

/* raw (unweighted, unscaled) E_g Jacobian rows [n_eg][29], same order as i3do_get_rows(type 0) */
int i3do_get_eg_jacobian(void* h, double* jac)
{
    auto* o = static_cast<Oracle*>(h);
    std::memcpy(jac, o->eg_jac.data(), sizeof(double) * o->eg_jac.size());
    return 0;
}

/* observation selection: [n][K] frames (-1 = none) and float weights, descending priority; active flags [n] */
int i3do_get_observations(void* h, int K, int32_t* frames, float* weights, uint8_t* active)
{
    auto* o = static_cast<Oracle*>(h);
    if (static_cast<size_t>(K) * o->n != o->obs_frame.size()) { o->error = "oracle: K mismatch"; return 1; }
    if (frames) std::memcpy(frames, o->obs_frame.data(), sizeof(int32_t) * o->obs_frame.size());
    if (weights) std::memcpy(weights, o->obs_weight.data(), sizeof(float) * o->obs_weight.size());
    if (active) std::memcpy(active, o->active.data(), o->n);
    return 0;
}

/* unknown-space vectors of the last iteration: layout [sdf n | albedo n | poses 6F | intr 4 | dist 5] */
int i3do_get_step(void* h, double* step, uint8_t* free_mask, double* col_scale)
{
    auto* o = static_cast<Oracle*>(h);
    const size_t U = static_cast<size_t>(o->num_unknowns());
    if (step && o->last_step.size() == U) std::memcpy(step, o->last_step.data(), sizeof(double) * U);
    if (free_mask && o->free_mask.size() == U) std::memcpy(free_mask, o->free_mask.data(), U);
    if (col_scale && o->col_scale.size() == U) std::memcpy(col_scale, o->col_scale.data(), sizeof(double) * U);
    return 0;
}

/* standalone evaluation of one E_g residual + raw Jacobian (KA1/KA2 tests) */
int i3do_eval_eg(const int32_t coord[3], double voxel_size, double pyr_scale, int w, int h, const float* lum, const double sh[9],
                 const double sdf[10], const double alb[4], const double pose[6], const double intr[4], const double dist[5],
                 double* residual, double* jac29)
{
    EgContext c; c.coord[0] = coord[0]; c.coord[1] = coord[1]; c.coord[2] = coord[2];
    c.voxel_size = voxel_size; c.pyr_scale = pyr_scale; c.w = w; c.h = h; c.lum = lum; c.sh = sh;
    *residual = eg_functor<double>(c, sdf, alb, pose, intr, dist);
    if (jac29)
    {
        for (int base = 0; base < I3D_EG_COLS; base += kStride)
        {
            Jet jsdf[10], jalb[4], jpose[6], jintr[4], jdist[5];
            for (int k = 0; k < 10; ++k) jsdf[k] = Jet(sdf[k]);
            for (int k = 0; k < 4; ++k) jalb[k] = Jet(alb[k]);
            for (int k = 0; k < 6; ++k) jpose[k] = Jet(pose[k]);
            for (int k = 0; k < 4; ++k) jintr[k] = Jet(intr[k]);
            for (int k = 0; k < 5; ++k) jdist[k] = Jet(dist[k]);
            for (int l = 0; l < kStride && base + l < I3D_EG_COLS; ++l)
            {
                const int p = base + l;
                if (p < 10) jsdf[p].v[l] = 1.0; else if (p < 14) jalb[p - 10].v[l] = 1.0;
                else if (p < 20) jpose[p - 14].v[l] = 1.0; else if (p < 24) jintr[p - 20].v[l] = 1.0; else jdist[p - 24].v[l] = 1.0;
            }
            const Jet res = eg_functor<Jet>(c, jsdf, jalb, jpose, jintr, jdist);
            for (int l = 0; l < kStride && base + l < I3D_EG_COLS; ++l) jac29[base + l] = res.v[l];
        }
    }
    return 0;
}

} // extern "C"
