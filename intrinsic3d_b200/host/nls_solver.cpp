// NLSSolver bound to the GPU engine (see include/nv/refinement/nls_solver.h).  Reference: src/refinement/nls_solver.cpp.
#include <nv/refinement/nls_solver.h>

#include <cstring>
#include <iostream>
#include <sstream>

#include "../../include/i3d_c_api.h"

namespace nv
{
std::string NLSSolver::ProblemInfo::toString(bool print_costs) const
{
    std::ostringstream o;
    o << "problem " << iteration << ": residuals " << residuals << ", parameters " << parameters;
    if (print_costs) o << ", cost " << cost;
    o << ", types [";
    for (size_t t = 0; t < type_residuals.size(); ++t) o << (t ? " " : "") << type_residuals[t];
    o << "], add " << time_add << " s, build " << time_build << " s";
    return o.str();
}
std::string NLSSolver::SolverInfo::toString() const
{
    std::ostringstream o;
    o << "solve " << iteration << ": cost " << cost << " -> " << cost_final << " (change " << cost_change << "), inner iterations " << inner_iterations
      << ", trust region radius " << trust_region_radius << ", " << time_solve << " s";
    return o.str();
}

NLSSolver::NLSSolver() { reset(1); }
NLSSolver::~NLSSolver() {}

bool NLSSolver::reset(size_t num_cost_types)
{
    if (num_cost_types == 0) return false;
    num_cost_types_ = num_cost_types;
    cost_type_weights_.assign(num_cost_types, 1.0);
    recorded_.assign(num_cost_types, 0);
    built_ = false; fix_poses_ = fix_intr_ = fix_dist_ = false;
    return true;
}

void NLSSolver::setCostWeight(size_t id, double w) { if (id < cost_type_weights_.size()) cost_type_weights_[id] = w; }
double NLSSolver::costWeight(size_t id) { return id < cost_type_weights_.size() ? cost_type_weights_[id] : 0.0; }
void NLSSolver::setDebug(bool d) { debug_ = d; }
void NLSSolver::attach(const Binding& b) { bind_ = b; }

bool NLSSolver::addResidual(VoxelResidual& r) { return addResidual(0, r); }

bool NLSSolver::addResidual(size_t cost_id, const VoxelResidual& r)
{
    if (cost_id >= cost_type_weights_.size()) return false;
    // ownership of the descriptor passes to the solver, like ceres::Problem takes the cost function
    const bool ok = r.cost != nullptr && r.weight != 0.0;
    if (ok && cost_id < 4 && static_cast<size_t>(r.cost->type) != cost_id)
    {
        std::cerr << "NLSSolver::addResidual: cost id " << cost_id << " does not match the built-in term type " << r.cost->type << std::endl;
        delete r.cost;
        return false;
    }
    delete r.cost;
    if (!ok) return false;
    recorded_[cost_id]++;
    return true;
}

namespace
{
I3DParams make_params(const NLSSolver::Binding& b, const std::vector<double>& w, bool fp, bool fi, bool fd, int lm_steps, bool build_only)
{
    I3DParams p;
    i3d_default_params(&p);
    for (size_t t = 0; t < 4; ++t) p.lambda[t] = t < w.size() ? w[t] : 0.0;
    p.use_er = b.use_er; p.use_es = b.use_es; p.use_ea = b.use_ea; p.fix_all_albedo = b.fix_all_albedo;
    p.thres_shell = b.thres_shell; p.occlusion_distance = b.occlusion_distance; p.num_observations = b.num_observations;
    p.lm_steps = lm_steps; p.fix_poses = fp; p.fix_intrinsics = fi; p.fix_distortion = fd; p.build_only = build_only;
    return p;
}
} // namespace

bool NLSSolver::buildProblem(bool use_normalized_weights)
{
    if (!bind_.engine) { std::cerr << "NLSSolver::buildProblem: no engine attached" << std::endl; return false; }
    if (!use_normalized_weights)
    {
        // the reference's Optimizer always normalises (optimizer.cpp:289); raw weights are not offered by the engine
        std::cerr << "NLSSolver::buildProblem: only use_normalized_weights = true is supported by the GPU engine" << std::endl;
        return false;
    }
    if (num_cost_types_ != 4) { std::cerr << "NLSSolver::buildProblem: the engine implements exactly the 4 built-in cost types" << std::endl; return false; }
    // Residual collection, weight normalisation and the Jacobian build all happen on the device inside solve() (one
    // i3d_gn_iteration call); ProblemInfo is filled from the same call so that the problem is not built twice.
    //
    // Contract for callers that DID record residuals through addResidual(): the engine always solves the complete four-term
    // problem of the attached grid (every residual Optimizer::addVoxelResiduals would add).  A recorded set that differs from
    // that enumeration cannot be honoured, so it is rejected here instead of silently solving a different problem: one
    // build-only pass on the device counts the engine's rows per type and buildProblem() fails on any mismatch.
    bool any_recorded = false;
    for (size_t t = 0; t < recorded_.size(); ++t) any_recorded = any_recorded || recorded_[t] != 0;
    if (any_recorded)
    {
        I3DParams p = make_params(bind_, cost_type_weights_, fix_poses_, fix_intr_, fix_dist_, 1, true);
        I3DIterInfo info;
        if (i3d_gn_iteration(bind_.engine, &p, &info) != 0) { std::cerr << "NLSSolver::buildProblem: " << i3d_last_error(bind_.engine) << std::endl; return false; }
        for (int t = 0; t < 4; ++t)
            if (recorded_[t] != static_cast<size_t>(info.type_residuals[t]))
            {
                std::cerr << "NLSSolver::buildProblem: " << recorded_[t] << " residuals of cost type " << t << " were added, but the engine's enumeration of the attached grid has "
                          << info.type_residuals[t] << "; subsets / custom residual sets are not supported (the engine solves the complete problem)" << std::endl;
                return false;
            }
    }
    built_ = true;
    return true;
}

bool NLSSolver::fixParamBlock(double* ptr)
{
    if (!built_ || !ptr) return false;
    if (bind_.intrinsics && ptr == bind_.intrinsics) { fix_intr_ = true; return true; }
    if (bind_.distortion && ptr == bind_.distortion) { fix_dist_ = true; return true; }
    if (bind_.poses_begin && ptr >= bind_.poses_begin && ptr < bind_.poses_begin + 6 * bind_.num_poses) { fix_poses_ = true; return true; }   // all poses or none
    // voxel blocks: the engine applies Optimizer::fixVoxelParams' rule itself (outside shell / invalid 1-ring / lambda_a < 0)
    return true;
}

bool NLSSolver::solve(int lm_steps)
{
    if (!bind_.engine || !built_) return false;
    I3DParams p = make_params(bind_, cost_type_weights_, fix_poses_, fix_intr_, fix_dist_, lm_steps, false);
    I3DIterInfo info;
    if (i3d_gn_iteration(bind_.engine, &p, &info) != 0) { std::cerr << "NLSSolver::solve: " << i3d_last_error(bind_.engine) << std::endl; return false; }
    {
        ProblemInfo pi;
        pi.iteration = problem_info_.size(); pi.residual_types = 4;
        for (int t = 0; t < 4; ++t)
        {
            pi.type_residuals.push_back(static_cast<size_t>(info.type_residuals[t]));
            pi.type_costs.push_back(info.type_costs[t]); pi.type_weights.push_back(info.type_weights[t]);
            pi.residuals += static_cast<size_t>(info.type_residuals[t]);
        }
        pi.parameters = static_cast<size_t>(info.num_parameters); pi.cost = info.cost_initial;
        pi.time_add = info.time_add; pi.time_build = info.time_build;
        problem_info_.push_back(pi);
        if (debug_) std::cout << "      " << pi.toString() << std::endl;
    }
    if (info.termination == 0 || info.termination == 1)
    {
        SolverInfo si;
        si.iteration = solver_info_.size(); si.cost = info.cost_initial; si.cost_final = info.cost_final; si.cost_change = si.cost - si.cost_final;
        si.inner_iterations = static_cast<size_t>(info.lm_iterations) + 1; si.trust_region_radius = info.trust_region_radius;
        si.time_solve = info.time_solve;
        std::ostringstream o;
        o << "LM trials " << info.lm_iterations << ", PCG iterations " << info.cg_iterations_total << ", accepted " << info.step_accepted;
        si.report = o.str();
        solver_info_.push_back(si);
        if (debug_) std::cout << "      " << si.toString() << std::endl;
    }
    built_ = false;
    return true;
}
} // namespace nv
