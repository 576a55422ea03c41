// nv/refinement/surface_stab_regularizer.h — SurfaceStabRegularizer plugin surface (reference: include/nv/refinement/surface_stab_regularizer.h,
// src/refinement/surface_stab_regularizer.cpp).  The row itself is evaluated matrix-free on the GPU (k_reg_build / k_op_partial).
#pragma once
#include <nv/refinement/cost.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class SurfaceStabRegularizer
{
public:
    static VoxelResidual create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v_pos);
};
} // namespace nv
