// nv/sdf/colorization.h — the slice of SDFColorization the optimiser uses (reference: include/nv/sdf/colorization.h:57-120):
// the Config with the occlusion distance and the number of best observations.  Observation selection itself
// (collectObservations, src/sdf/colorization.cpp:192-370) runs on the GPU inside i3d_gn_iteration.
#pragma once
#include <cstddef>

#include <nv/mat.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class SDFColorization
{
public:
    struct Config
    {
        int discont_distance = 0;
        Vec3b color_unobserved = Vec3b::Zero();
        float color_range = 20.0f;
        float max_occlusion_distance = 0.05f;
        size_t max_num_observations = 5;
    };
    SDFColorization() = default;
    explicit SDFColorization(SparseVoxelGrid<VoxelSBR>* grid) : grid_(grid) {}
    void setConfig(const Config& cfg) { cfg_ = cfg; }
    const Config& config() const { return cfg_; }
    bool reset(SparseVoxelGrid<VoxelSBR>* grid, const Vec4& intrinsics, const Vec5& dist, int w, int h)
    {
        grid_ = grid; intrinsics_ = intrinsics; dist_ = dist; w_ = w; h_ = h;
        return grid != nullptr && !grid->empty();
    }

private:
    Config cfg_;
    SparseVoxelGrid<VoxelSBR>* grid_ = nullptr;
    Vec4 intrinsics_;
    Vec5 dist_;
    int w_ = 0, h_ = 0;
};
} // namespace nv
