// nv/refinement/cost.h — residual descriptor + scalar helpers (reference: include/nv/refinement/cost.h:59-150).
// In the reference a VoxelResidual carries a heap-allocated ceres::CostFunction; here `cost` is a small tagged
// descriptor: the arithmetic of all four built-in terms lives in the GPU engine, so a descriptor only records WHICH row
// the caller asked for.  The "not applicable" conventions are kept: cost == nullptr or weight == 0.
#pragma once
#include <cmath>
#include <vector>

#include <nv/mat.h>

#define NV_INVALID_RESIDUAL 0.0

namespace nv
{
struct CostTerm
{
    enum Type { SHADING = 0, VOLUMETRIC = 1, SURFACE_STAB = 2, ALBEDO = 3 } type;
    Vec3i v_pos;       // owning voxel
    Vec3i v_pos_nb;    // albedo regulariser: neighbour
    const void* frame; // shading: ShadingCostData of the observing frame
};

struct VoxelResidual
{
    CostTerm* cost = nullptr;
    std::vector<double*> params;
    double weight = 0.0;
};

// linear ramp of a cost weight over the outer iterations (cost.h:130-143)
inline double computeVaryingLambda(int iteration, int num_iterations, double lambda0, double lambda1)
{
    if (num_iterations <= 1) return lambda0;
    return lambda0 + (lambda1 - lambda0) / static_cast<double>(num_iterations - 1) * static_cast<double>(iteration);
}
inline double pyramidLevelToScale(int lvl) { return 1.0 / std::pow(2.0, lvl); }
} // namespace nv
