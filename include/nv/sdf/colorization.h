// nv/sdf/colorization.h — SDFColorization with the reference's API surface for the refinement path
// (include/nv/sdf/colorization.h:57-120): the Config the optimiser reads (occlusion distance, number of best observations),
// reset(), and the recolouring pair add() / compute() that Intrinsic3D::recomputeColors drives
// (src/refinement/intrinsic3d.cpp:381-409, src/sdf/colorization.cpp:113-189).
//
// Observation selection for the optimiser (collectObservations, colorization.cpp:192-370) runs on the GPU inside
// i3d_gn_iteration.  add() here only records the view (the reference computes that view's observations immediately and keeps
// N x F VertexObservation objects); compute() runs ONE device pass over all recorded views (i3d_recompute_colors) and writes
// VoxelSBR::color.  The grid must not change between add() and compute() (it does not in the reference's only caller).
#pragma once
#include <cstddef>
#include <vector>

#include <nv/image.h>
#include <nv/mat.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class SDFColorization
{
public:
    struct Config
    {
        int discont_distance = 0;          // erodeDiscontinuities radius: its result is dead code in add() (colorization.cpp:126,146)
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
        views_.clear();
        return grid != nullptr && !grid->empty();
    }
    // colorization.cpp:113-158
    bool add(int id, const ImageF& depth, const ImageBGR& color, const Mat4f& pose_world_to_cam);
    // colorization.cpp:161-189
    bool compute();
    void setDevice(int cuda_device) { device_ = cuda_device; }

private:
    struct View { int id; ImageF depth; ImageBGR color; Mat4f pose; };
    Config cfg_;
    SparseVoxelGrid<VoxelSBR>* grid_ = nullptr;
    Vec4 intrinsics_;
    Vec5 dist_;
    int w_ = 0, h_ = 0;
    int device_ = 0;
    std::vector<View> views_;
};
} // namespace nv
