// nv/refinement/shading_cost.h — E_g data-term plugin surface (reference: include/nv/refinement/shading_cost.h:52-83,
// src/refinement/shading_cost.cpp:59-150).  create() keeps the reference's signature and applicability tests that do not need
// image access (stencil existence, valid normal); the residual value / Jacobian are produced by the GPU kernels
// (intrinsic3d_b200/csrc/i3d_math.cuh), which also drop rows whose residual evaluates to the 0.0 sentinel.
#pragma once
#include <nv/refinement/cost.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class ShadingCostData
{
public:
    ShadingCostData(int rgbd_lvl, double vx_size, int w, int h, const float* ptr_lum)
        : pyr_scale(pyramidLevelToScale(rgbd_lvl)), voxel_size(vx_size), width(w), height(h), ptr_intensity(ptr_lum) {}
    double pyr_scale;
    double voxel_size;
    int width, height;
    const float* ptr_intensity;
};

class ShadingCost
{
public:
    static VoxelResidual create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v_pos, Vec6& pose_vec, Vec4& intrinsics, Vec5& distortion_coeffs,
                                const VecXd& sh_coeffs, const ShadingCostData* data);
};
} // namespace nv
