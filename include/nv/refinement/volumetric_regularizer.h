// nv/refinement/volumetric_regularizer.h — VolumetricRegularizer plugin surface (reference: include/nv/refinement/volumetric_regularizer.h,
// src/refinement/volumetric_regularizer.cpp).  The row itself is evaluated matrix-free on the GPU (k_reg_build / k_op_partial).
#pragma once
#include <nv/refinement/cost.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class VolumetricRegularizer
{
public:
    static VoxelResidual create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v_pos);
};
} // namespace nv
