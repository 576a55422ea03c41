// Cost-term plugin entry points with the reference's create() signatures.  They perform the structural applicability
// tests of the reference (which voxels must exist / be valid) and return a descriptor; residuals and Jacobians are
// evaluated by the GPU engine.  Reference: src/refinement/shading_cost.cpp:59-150, volumetric_regularizer.cpp:52-78,
// surface_stab_regularizer.cpp:51-61, albedo_regularizer.cpp:50-84.
#include <cmath>

#include <nv/refinement/albedo_regularizer.h>
#include <nv/refinement/shading_cost.h>
#include <nv/refinement/surface_stab_regularizer.h>
#include <nv/refinement/volumetric_regularizer.h>

namespace nv
{
namespace
{
bool has_normal(const SparseVoxelGrid<VoxelSBR>* g, const Vec3i& p)
{
    // float forward differences need v, +x, +y, +z valid and a non-zero gradient (src/sdf/operators.cpp:58-77)
    const Vec3i px{p[0] + 1, p[1], p[2]}, py{p[0], p[1] + 1, p[2]}, pz{p[0], p[1], p[2] + 1};
    if (!g->valid(p) || !g->valid(px) || !g->valid(py) || !g->valid(pz)) return false;
    const float s0 = static_cast<float>(g->voxel(p).sdf_refined);
    const float gx = static_cast<float>(g->voxel(px).sdf_refined) - s0;
    const float gy = static_cast<float>(g->voxel(py).sdf_refined) - s0;
    const float gz = static_cast<float>(g->voxel(pz).sdf_refined) - s0;
    return !(gx == 0.0f && gy == 0.0f && gz == 0.0f);
}
bool ring_valid(const SparseVoxelGrid<VoxelSBR>* g, const Vec3i& p)
{
    static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (const auto& o : d) if (!g->valid(p[0] + o[0], p[1] + o[1], p[2] + o[2])) return false;
    return true;
}
} // namespace

VoxelResidual ShadingCost::create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v, Vec6& pose_vec, Vec4& intrinsics, Vec5& dist,
                                  const VecXd& sh_coeffs, const ShadingCostData* data)
{
    VoxelResidual r;
    if (!grid || !data || sh_coeffs.size() < 9) return r;
    static const int outer[6][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {0, 1, 1}, {1, 1, 0}, {1, 0, 1}};
    for (const auto& o : outer) if (!grid->exists(v[0] + o[0], v[1] + o[1], v[2] + o[2])) return r;
    if (!has_normal(grid, v)) return r;
    r.weight = 1.0;
    r.cost = new CostTerm{CostTerm::SHADING, v, v, data};
    // the 17 parameter blocks in the reference's order: 10 sdf, 4 albedo, pose, intrinsics, distortion
    static const int so[10][3] = {{0, 0, 0}, {0, 1, 0}, {0, 2, 0}, {0, 1, 1}, {0, 0, 1}, {0, 0, 2}, {1, 0, 0}, {1, 1, 0}, {1, 0, 1}, {2, 0, 0}};
    for (const auto& o : so) r.params.push_back(&grid->voxel(v[0] + o[0], v[1] + o[1], v[2] + o[2]).sdf_refined);
    static const int ao[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (const auto& o : ao) r.params.push_back(&grid->voxel(v[0] + o[0], v[1] + o[1], v[2] + o[2]).albedo);
    r.params.push_back(pose_vec.data());
    r.params.push_back(intrinsics.data());
    r.params.push_back(dist.data());
    return r;
}

VoxelResidual VolumetricRegularizer::create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v)
{
    VoxelResidual r;
    if (!grid || !ring_valid(grid, v)) return r;
    r.cost = new CostTerm{CostTerm::VOLUMETRIC, v, v, nullptr};
    r.weight = 1.0;
    r.params.push_back(&grid->voxel(v).sdf_refined);
    static const int d[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (const auto& o : d) r.params.push_back(&grid->voxel(v[0] + o[0], v[1] + o[1], v[2] + o[2]).sdf_refined);
    return r;
}

VoxelResidual SurfaceStabRegularizer::create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v)
{
    VoxelResidual r;
    if (!grid || !grid->exists(v)) return r;
    r.cost = new CostTerm{CostTerm::SURFACE_STAB, v, v, nullptr};
    r.weight = 1.0;
    r.params.push_back(&grid->voxel(v).sdf_refined);
    return r;
}

VoxelResidual AlbedoRegularizer::create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v, const Vec3i& nb)
{
    VoxelResidual r;
    if (!grid || !grid->valid(v) || !grid->valid(nb)) return r;
    VoxelSBR& a = grid->voxel(v);
    VoxelSBR& b = grid->voxel(nb);
    auto lum = [](const Vec3b& c) { return 0.299f * c[0] + 0.587f * c[1] + 0.114f * c[2]; };
    const float la = lum(a.color), lb = lum(b.color);
    float d2 = 0.0f;
    for (int k = 0; k < 3; ++k)
    {
        const float d = (a.color[k] * (1.0f / 255.0f)) / la - (b.color[k] * (1.0f / 255.0f)) / lb;
        d2 += d * d;
    }
    const float t = 1.0f - std::sqrt(d2);
    const float chroma = (t < 0.01f) ? 0.01f : t;
    const double w = static_cast<double>(chroma);
    if (std::isnan(w) || std::isinf(w)) return r;
    r.cost = new CostTerm{CostTerm::ALBEDO, v, nb, nullptr};
    r.weight = w;
    r.params.push_back(&a.albedo);
    r.params.push_back(&b.albedo);
    return r;
}
} // namespace nv
