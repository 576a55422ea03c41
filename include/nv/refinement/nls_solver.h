// nv/refinement/nls_solver.h — NLSSolver with the reference's method surface (include/nv/refinement/nls_solver.h:53-125),
// bound to the GPU engine instead of ceres::Problem / ceres::Solve.
//
// Call sequence kept from Optimizer::optimize: reset(4) -> setCostWeight(0..3) -> [addResidual ...] -> buildProblem(true)
// -> fixParamBlock(...) -> solve(lm_steps).  Differences that follow from moving the arithmetic to the GPU:
//   * the four built-in cost types (ids 0..3 = E_g, E_r, E_s, E_a) are the only ones supported, and the engine ALWAYS solves the
//     complete problem of the attached grid: it enumerates every residual itself, exactly as Optimizer::addVoxelResiduals would.
//     addResidual() keeps the reference's ownership and return conventions and records the per-type counts; if anything was
//     recorded, buildProblem() verifies the counts against the engine's enumeration (one build-only pass on the device) and returns
//     false on a mismatch — a caller-chosen SUBSET of residuals is rejected, never silently replaced by the full problem
//     (reference contract: src/refinement/nls_solver.cpp:172-187);
//   * fixParamBlock() recognises the camera blocks (poses / intrinsics / distortion); voxel parameters are fixed by the
//     engine with the rule of Optimizer::fixVoxelParams.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include <nv/refinement/cost.h>

struct I3DEngine;

namespace nv
{
class NLSSolver
{
public:
    struct ProblemInfo
    {
        size_t iteration = 0, residuals = 0, parameters = 0;
        double cost = 0.0;
        size_t residual_types = 0;
        std::vector<size_t> type_residuals;
        std::vector<double> type_costs, type_weights;
        double time_add = 0.0, time_build = 0.0;
        std::string toString(bool print_costs = true) const;
    };
    struct SolverInfo
    {
        size_t iteration = 0;
        double cost = 0.0, cost_final = 0.0, cost_change = 0.0;
        size_t inner_iterations = 0;
        double trust_region_radius = 0.0;
        std::string report;
        double time_solve = 0.0;
        std::string toString() const;
    };

    NLSSolver();
    ~NLSSolver();

    bool reset(size_t num_cost_types = 1);
    bool addResidual(VoxelResidual& residual);
    bool addResidual(size_t cost_id, const VoxelResidual& residual);
    void setCostWeight(size_t cost_id, double weight);
    double costWeight(size_t cost_id);
    void setDebug(bool debug);
    bool buildProblem(bool use_normalized_weights = false);
    bool solve(int lm_steps);
    bool fixParamBlock(double* ptr);

    // ---- binding to the engine (not in the reference) ----
    struct Binding
    {
        I3DEngine* engine = nullptr;
        double thres_shell = 0.0;
        float occlusion_distance = 0.02f;
        int num_observations = 5;
        bool use_er = true, use_es = true, use_ea = true, fix_all_albedo = false;
        const double* poses_begin = nullptr; size_t num_poses = 0;     // to recognise camera blocks in fixParamBlock
        const double* intrinsics = nullptr; const double* distortion = nullptr;
    };
    void attach(const Binding& b);
    const std::vector<ProblemInfo>& problemInfo() const { return problem_info_; }
    const std::vector<SolverInfo>& solverInfo() const { return solver_info_; }

private:
    size_t num_cost_types_ = 1;
    std::vector<double> cost_type_weights_;
    std::vector<size_t> recorded_;
    std::vector<ProblemInfo> problem_info_;
    std::vector<SolverInfo> solver_info_;
    Binding bind_;
    bool fix_poses_ = false, fix_intr_ = false, fix_dist_ = false;
    bool built_ = false, debug_ = false;
};
} // namespace nv
