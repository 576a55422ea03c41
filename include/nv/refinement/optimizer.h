// nv/refinement/optimizer.h — Optimizer with the reference's API surface (include/nv/refinement/optimizer.h:59-141), driving the
// B200 engine through the C-ABI (include/i3d_c_api.h) instead of Ceres.
//
// optimize() mutates, in place and like the reference: grid voxels' sdf_refined / albedo, image_formation.poses /
// intrinsics / distortion_coeffs.  Returns false only if the grid is null or iterations < 1 (optimizer.cpp:113-114) or if the
// engine reports an error (message on std::cerr).
#pragma once
#include <map>
#include <string>
#include <unordered_set>
#include <vector>

#include <nv/mat.h>
#include <nv/refinement/nls_solver.h>
#include <nv/refinement/shading_cost.h>
#include <nv/rgbd/pyramid.h>
#include <nv/sdf/colorization.h>
#include <nv/sparse_voxel_grid.h>

namespace nv
{
class Optimizer
{
public:
    struct Config
    {
        int iterations = 10;
        int lm_steps = 50;
        double lambda_g = 0.2;
        double lambda_r0 = 20.0, lambda_r1 = 160.0;
        double lambda_s0 = 10.0, lambda_s1 = 120.0;
        double lambda_a = 0.1;
        bool fix_poses = false, fix_intrinsics = false, fix_distortion = false;
        // flat key -> value settings with the key names of data/intrinsic3d.yml (the reference reads them through
        // nv::Settings / cv::FileStorage, src/refinement/optimizer.cpp:52-72); missing keys keep the defaults above
        void load(const std::map<std::string, std::string>& settings);
        void print() const;
    };
    struct Data
    {
        SparseVoxelGrid<VoxelSBR>* grid = nullptr;     // not owned
        double thres_shell = 0.0;
        int grid_level = 0;
        int rgbd_level = 0;
        std::vector<VecXd> voxel_sh_coeffs;            // indexed by the grid's iteration order
        std::vector<ShadingCostData> shading_cost_data;
        std::unordered_set<Vec3i, std::hash<Vec3i>> voxels_added;   // scratch in the reference; unused here
    };
    struct ImageFormationModel
    {
        Vec4 intrinsics = Vec4::Zero();
        Vec5 distortion_coeffs = Vec5::Zero();
        std::vector<int> frame_ids;
        std::vector<Vec6> poses;
        std::vector<Pyramid> rgbd_pyr;
    };

    explicit Optimizer(Config cfg);
    ~Optimizer();
    const Config& config() const;
    bool optimize(SDFColorization& colorization, Data& data, ImageFormationModel& image_formation);

    // per outer iteration diagnostics of the last optimize() call (NLSSolver::ProblemInfo / SolverInfo)
    const std::vector<NLSSolver::ProblemInfo>& problemInfo() const { return problem_info_; }
    const std::vector<NLSSolver::SolverInfo>& solverInfo() const { return solver_info_; }
    void setDevice(int cuda_device) { device_ = cuda_device; }

private:
    Config cfg_;
    int device_ = 0;
    std::vector<NLSSolver::ProblemInfo> problem_info_;
    std::vector<NLSSolver::SolverInfo> solver_info_;
};
} // namespace nv
