/*
 * i3d_types.h — plain-C parameter / result structs shared by the B200 engine
 * C-ABI (include/i3d_c_api.h) and by the CPU oracle (oracle/oracle_api.h).
 *
 * Field meanings follow the reference (paths relative to the NVlabs/intrinsic3d
 * tree, libintrinsic3d/ = L/):
 *   - cost-type ids 0..3 = E_g, E_r, E_s, E_a as set in
 *     L/src/refinement/optimizer.cpp:140-143
 *   - solver options = Ceres 2.1.0 defaults the reference leaves untouched
 *     (L/src/refinement/nls_solver.cpp:300-337 sets only max_num_iterations,
 *     CGNR, num_threads).
 */
#ifndef I3D_TYPES_H_
#define I3D_TYPES_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I3D_NUM_COST_TYPES 4
#define I3D_MAX_OBS 8          /* upper bound for num_observations (reference default 5) */
#define I3D_EG_COLS 29         /* 10 sdf + 4 albedo + 6 pose + 4 intrinsics + 5 distortion */
#define I3D_MAX_LM_STEPS 64    /* per-trial bookkeeping slots in I3DIterInfo */

/* Parameters of ONE Gauss-Newton (outer) iteration of Optimizer::optimize
 * (L/src/refinement/optimizer.cpp:119-171). The host shim computes the ramped
 * lambdas (L/include/nv/refinement/cost.h:130-143) and the term switches. */
typedef struct I3DParams
{
    /* NLSSolver::setCostWeight(0..3): lambda_g, lambda_r(itr), lambda_s(itr), lambda_a */
    double lambda[I3D_NUM_COST_TYPES];
    /* term switches: E_r iff lambda_r0>0 && lambda_r1>0, E_s likewise, E_a iff lambda_a>0
     * (optimizer.cpp:239,249,259) */
    int32_t use_er, use_es, use_ea;
    /* lambda_a < 0  =>  every albedo is constant (optimizer.cpp:330-334) */
    int32_t fix_all_albedo;
    /* Optimizer::Data::thres_shell */
    double thres_shell;
    /* SDFColorization::Config::max_occlusion_distance (0 => no occlusion test), float like the reference */
    float occlusion_distance;
    /* SDFColorization::Config::max_num_observations (K); 0 => keep all frames */
    int32_t num_observations;
    /* Optimizer::Config::lm_steps  -> ceres max_num_iterations */
    int32_t lm_steps;
    int32_t fix_poses, fix_intrinsics, fix_distortion;

    /* ---- Ceres 2.1.0 defaults (restated; see oracle/oracle.cpp header) ---- */
    double initial_trust_region_radius;   /* 1e4  */
    double max_trust_region_radius;       /* 1e16 */
    double min_trust_region_radius;       /* 1e-32 */
    double min_relative_decrease;         /* 1e-3 */
    double min_lm_diagonal;               /* 1e-6 */
    double max_lm_diagonal;               /* 1e32 */
    double eta;                           /* 0.1  (CG q-tolerance) */
    double function_tolerance;            /* 1e-6 */
    double gradient_tolerance;            /* 1e-10 */
    double parameter_tolerance;           /* 1e-8 */
    int32_t max_linear_solver_iterations; /* 500 */
    int32_t min_linear_solver_iterations; /* 0 */
    int32_t residual_reset_period;        /* 10 */
    int32_t max_consecutive_invalid_steps;/* 5 */

    /* ---- test hooks (not in the reference) ---- */
    /* >0: every CGNR solve runs exactly this many iterations (parity tests compare
     * engine and oracle at an identical iteration count). */
    int32_t forced_cg_iterations;
    /* 1: evaluate and build the problem, but do not run the LM solve */
    int32_t build_only;
} I3DParams;

/* Mirrors NLSSolver::ProblemInfo + SolverInfo (L/include/nv/refinement/nls_solver.h:58-89)
 * plus what the parity tests need. */
typedef struct I3DIterInfo
{
    int64_t num_voxels;
    int64_t num_active;                       /* voxels passing addVoxelResiduals' tests */
    int64_t num_free_sdf, num_free_albedo;    /* by mask (fixVoxelParams) */
    int64_t num_parameters;                   /* free scalars incl. camera */
    int64_t type_residuals[I3D_NUM_COST_TYPES];
    double  type_sum_weights[I3D_NUM_COST_TYPES];  /* sum of raw residual weights per type */
    double  type_weights[I3D_NUM_COST_TYPES];      /* lambda/sum*1000 (ProblemInfo::type_weights) */
    double  type_costs[I3D_NUM_COST_TYPES];        /* 0.5*sum w r^2 at the initial point, per type */
    double  cost_initial;                     /* SolverInfo::cost        */
    double  cost_final;                       /* SolverInfo::cost_final  */
    double  trust_region_radius;              /* radius after the last LM iteration */
    int32_t lm_iterations;                    /* trial steps taken (SolverInfo::inner_iterations - 1) */
    int32_t step_accepted;                    /* 1 if a successful step was applied */
    int32_t termination;                      /* 0 user_success(first accepted step) 1 convergence
                                                 2 no_convergence 3 failure 4 nothing to do */
    int32_t cg_iterations_total;
    int32_t cg_iterations[I3D_MAX_LM_STEPS];  /* per trial */
    double  model_cost_change[I3D_MAX_LM_STEPS];
    double  candidate_cost[I3D_MAX_LM_STEPS];
    double  relative_decrease[I3D_MAX_LM_STEPS];
    double  step_norm;                        /* ||delta|| of the last evaluated trial (unscaled) */
    double  time_add, time_build, time_solve; /* seconds; the reference's three phase timers */
} I3DIterInfo;

/* ---- SVSH lighting (LightingSVSH, libintrinsic3d/src/lighting/lighting_svsh.cpp:54-346) ---- */
#define I3D_SH_COEFFS 9

/* Constructor arguments of LightingSVSH (include/nv/lighting/lighting_svsh.h:50) plus the Ceres
 * options LightingSVSH::estimate sets (lighting_svsh.cpp:186,325-337) and the Ceres 2.1.0
 * defaults it inherits. */
typedef struct I3DLightingParams
{
    float   subvolume_size;               /* Intrinsic3D::Config::subvolume_size_sh (0.2 m) */
    int32_t weighted;                     /* refine() passes true: weight = sdfToWeight(sdf_refined, truncation) */
    double  lambda_reg;                   /* sh_est_lambda_reg (10.0) */
    double  thres_shell;                  /* Optimizer::Data::thres_shell */
    int32_t max_iterations;               /* 50 (lighting_svsh.cpp:186) */
    int32_t max_linear_solver_iterations; /* 500 */
    int32_t min_linear_solver_iterations; /* 0 */
    int32_t residual_reset_period;        /* 10 */
    int32_t max_consecutive_invalid_steps;/* 5 */
    int32_t reserved;
    double  initial_trust_region_radius;  /* 1e4 */
    double  max_trust_region_radius;      /* 1e16 */
    double  min_trust_region_radius;      /* 1e-32 */
    double  min_relative_decrease;        /* 1e-3 */
    double  min_lm_diagonal;              /* 1e-6 */
    double  max_lm_diagonal;              /* 1e32 */
    double  eta;                          /* 0.1 */
    double  function_tolerance;           /* 1e-6 */
    double  gradient_tolerance;           /* 1e-10 */
    double  parameter_tolerance;          /* 1e-8 */
} I3DLightingParams;

/* What ceres::Solver::Summary would report for the SH problem, plus problem sizes. */
typedef struct I3DLightingInfo
{
    int64_t num_subvolumes;               /* Subvolumes::count() */
    int64_t num_data_rows;                /* SHDataCost residuals (one per contributing voxel) */
    int64_t num_reg_pairs;                /* SHRegularizerCost blocks (directed pairs, 9 residuals each) */
    double  sum_data_weights;
    double  cost_initial, cost_final;
    double  trust_region_radius;
    int32_t lm_iterations;                /* trust-region iterations started (Summary::iterations.size() - 1) */
    int32_t num_successful_steps;
    int32_t cg_iterations_total;
    int32_t termination;                  /* 0 convergence, 1 no_convergence (max iterations), 2 failure */
    int32_t usable;                       /* Summary::IsSolutionUsable() == the return value of estimate() */
    int32_t reserved;
    double  time_accumulate, time_solve, time_interpolate;   /* seconds (device time) */
} I3DLightingInfo;

#ifdef __cplusplus
}
#endif
#endif /* I3D_TYPES_H_ */
