// nv/refinement/albedo_regularizer.h — AlbedoRegularizer plugin surface (reference: include/nv/refinement/albedo_regularizer.h,
// src/refinement/albedo_regularizer.cpp:50-84): chromaticity-weighted albedo difference of two neighbouring voxels.
#pragma once
#include <nv/refinement/cost.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class AlbedoRegularizer
{
public:
    static VoxelResidual create(SparseVoxelGrid<VoxelSBR>* grid, const Vec3i& v_pos, const Vec3i& v_pos_nb);
};
} // namespace nv
